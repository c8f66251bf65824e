# Refresh of the judged numbers for the current kernels: GPU tests, smoke, bench lines, launch list, ncu --set full.
# Usage (from the repo root, on the GPU box): bash profiles/scripts/round_refresh.sh <tag>
tag=${1:-v3}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_lw_$tag.json 2> $out/bench_lw_$tag.err; echo "bench lw rc=$?"
for w in amars_sw gcm lbl; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $out/bench_${w}_$tag.json 2> $out/bench_${w}_$tag.err; echo "bench $w rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_$tag.json 2>&1; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_lw_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
for w in amars_lw amars_sw gcm lbl; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:disort_flux_kernel -s 3 -c 1 -f -o $out/prof_${w}_$tag \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-e2e > $out/ncu_full_${w}_$tag.log 2>&1; echo "ncu full $w rc=$?"
done
tail -3 $out/pytest_gpu_$tag.log
