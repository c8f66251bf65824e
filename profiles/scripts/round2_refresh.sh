# Refresh of the judged single-GPU numbers for the current kernels (round 2): GPU tests, smoke, ncu --set full of the solve
# kernel per workload and of the optics kernel, launch lists, traffic.json (stamped with the hash of the sources in the tree),
# THEN the bench lines (so that their roofline.frac uses the fresh counters), parity report, latency probe.
# Usage (repo root, on the GPU box): bash profiles/scripts/round2_refresh.sh <tag>
tag=${1:-v1}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_lw_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-others > $out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"disort_flux_kernel|band_optics" -s 6 -c 2 -f -o $out/prof_amars_lw_$tag \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-others > $out/ncu_full_amars_lw_$tag.log 2>&1; echo "ncu full amars_lw rc=$?"
for w in amars_sw lbl; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:disort_flux_kernel -s 3 -c 1 -f -o $out/prof_${w}_$tag \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-e2e > $out/ncu_full_${w}_$tag.log 2>&1; echo "ncu full $w rc=$?"
done
# nstr = 16 runs the layer-split form: 7 chunks x 3 kernels per step at 16384 columns (the chunks differ: other spectral
# points); skip the 3 warm-up steps; --set full of the first chunk, then the counters of traffic.json over the 21 launches of a step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"split_prefix|split_layer|team_adding" -s 63 -c 3 -f -o $out/prof_gcm_$tag \
  python bench.py --workload gcm --ncol 16384 --steps 2 --warmup 3 --no-cpu --no-e2e > $out/ncu_full_gcm_$tag.log 2>&1; echo "ncu full gcm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum \
  --clock-control none -k regex:"split_prefix|split_layer|team_adding" -s 63 -c 21 -f -o $out/prof_gcm_step_$tag \
  python bench.py --workload gcm --ncol 16384 --steps 2 --warmup 3 --no-cpu --no-e2e > $out/ncu_step_gcm_$tag.log 2>&1; echo "ncu step gcm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"split_|team_" -c 120 --csv --log-file $out/launches_gcm_$tag.csv \
  python bench.py --workload gcm --ncol 16384 --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1; echo "ncu launches gcm rc=$?"
python profiles/scripts/make_traffic.py $tag amars_lw=$out/prof_amars_lw_$tag.ncu-rep:524288 amars_sw=$out/prof_amars_sw_$tag.ncu-rep:512000 \
  lbl=$out/prof_lbl_$tag.ncu-rep:800000 gcm=$out/prof_gcm_step_$tag.ncu-rep:524288 > $out/make_traffic_$tag.log 2>&1; echo "traffic rc=$?"
cp profiles/r02/traffic.json $out/traffic.json
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench_lw_$tag.json 2> $out/bench_lw_$tag.err; echo "bench default rc=$?"
for w in amars_sw gcm lbl; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > $out/bench_${w}_$tag.json 2> $out/bench_${w}_$tag.err; echo "bench $w rc=$?"
done
timeout 600 python bench.py --workload lbl --shard wave --steps 50 --warmup 5 > $out/bench_lblwave_1gpu_$tag.json 2> $out/bench_lblwave_1gpu_$tag.err; echo "bench lbl wave rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_$tag.json 2>&1; echo "bench ref rc=$?"
timeout 300 python profiles/scripts/latency_probe.py > $out/latency_$tag.txt 2>&1; echo "latency rc=$?"
timeout 600 python profiles/scripts/parity_report.py > $out/parity_report_$tag.txt 2>&1; echo "parity rc=$?"
tail -3 $out/pytest_gpu_$tag.log
