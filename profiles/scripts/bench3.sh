# the three throughput workloads, kernel-only (no cpu/e2e legs)
for w in amars_lw amars_sw gcm lbl; do
  echo -n "$w: "
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms frac %.3f'%(d['value'], d['kernel']['avg_ms'], d['roofline']['frac']))"
done
