for lib in libb200disort_e1.0e-31.so libb200disort.so libb200disort_e1.0e-22.so; do
  for w in amars_sw gcm; do
    echo -n "$lib $w: "
    B200DISORT_LIB=$PWD/pyharp_b200/$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms'%(d['value'], d['kernel']['avg_ms']))"
  done
done
