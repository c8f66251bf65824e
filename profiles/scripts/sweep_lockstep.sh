for b in 128 256; do
  for w in amars_lw amars_sw gcm; do
    echo -n "block=$b $w: "
    B200DISORT_BLOCK=$b python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms'%(d['value'], d['kernel']['avg_ms']))"
  done
done
