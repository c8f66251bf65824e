# Kernel experiments: builds pyharp_b200/libb200disort_<tag>.so = the product library with the nn = <nn> flux
# translation unit recompiled under extra -D switches.  Usage: bash profiles/scripts/mk_variant.sh <tag> <nn> "<-D...>"
tag=$1; nn=$2; extra=$3
cd "$(dirname "$0")/../../pyharp_b200/csrc"
mkdir -p ../../build_e40
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -DB200_NN=$nn -DB200_UNROLL_MAX_NN=12 -DB200_SPLIT_UNROLL_MAX_NN=16 -DB200_MIN_BLOCKS=0 $extra \
  -c flux_inst.cu -o ../../build_e40/flux_nn${nn}_$tag.o 2> ../../build_e40/ptxas_nn${nn}_$tag.log || { cat ../../build_e40/ptxas_nn${nn}_$tag.log; exit 1; }
objs=$(ls build/*.o | grep -v "flux_nn${nn}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libb200disort_$tag.so $objs ../../build_e40/flux_nn${nn}_$tag.o -ldl
grep -A2 "split_layer\|team_adding_kernelILi1" ../../build_e40/ptxas_nn${nn}_$tag.log | grep -E "stack|registers"
