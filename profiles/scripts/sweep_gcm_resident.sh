# GCM (nstr=16): resident threads per SM vs the L2 capacity for the 3.2 KB/thread spill frame
for cfg in "128 0" "96 2" "64 3" "128 1" "256 1" "192 1"; do
  set -- $cfg
  echo -n "block $1 ctas/SM ${2}: "
  B200DISORT_BLOCK=$1 B200DISORT_CTAS_PER_SM=$2 python bench.py --workload gcm --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms'%(d['value'], d['kernel']['avg_ms']))"
done
