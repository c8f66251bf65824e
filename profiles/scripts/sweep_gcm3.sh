for lib in libb200disort.so libb200disort_mb3.so libb200disort_mb4.so; do
  echo -n "$lib gcm: "
  B200DISORT_LIB=$PWD/pyharp_b200/$lib python bench.py --workload gcm --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms'%(d['value'], d['kernel']['avg_ms']))"
done
