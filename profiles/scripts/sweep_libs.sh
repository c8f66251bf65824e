# A/B of kernel variants built as pyharp_b200/libb200disort_<tag>.so (make BUILD=build_e_<tag> OUT=../libb200disort_<tag>.so EXTRA=...)
# usage: bash profiles/scripts/sweep_libs.sh "<tags>" "<workloads>"   ('' = the default library)
tags=${1:-"default"}
wls=${2:-"amars_lw amars_sw gcm lbl"}
for t in $tags; do
  lib=$PWD/pyharp_b200/libb200disort_$t.so; [ "$t" = default ] && lib=$PWD/pyharp_b200/libb200disort.so
  for w in $wls; do
    echo -n "$t $w: "
    B200DISORT_LIB=$lib python bench.py --workload $w --steps 8 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms frac %.3f'%(d['value'], d['kernel']['avg_ms'], d['roofline']['frac']))"
  done
done
