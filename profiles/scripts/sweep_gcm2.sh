for cfg in "128 1" "128 2" "64 2" "64 3" "64 4" "96 2"; do set -- $cfg
  echo -n "block=$1 ctas=$2: "
  B200DISORT_BLOCK=$1 B200DISORT_CTAS_PER_SM=$2 python bench.py --workload gcm --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3g solves/s kernel %.2f ms'%(d['value'], d['kernel']['avg_ms']))"
done
