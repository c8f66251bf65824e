for b in 32 64 128; do for c in 2 4 8; do
  echo -n "block=$b ctas=$c: "
  B200DISORT_BLOCK=$b B200DISORT_CTAS_PER_SM=$c python bench.py --workload gcm --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3g solves/s kernel %.2f ms'%(d['value'], d['kernel']['avg_ms']))"
done; done
