# GCM (nstr=16): resident threads per SM vs the L2 capacity for the 3.2 KB/thread spill frame
# usage: CFGS="block:ctas ..." bash profiles/scripts/sweep_gcm_resident.sh     (ctas 0 = whatever fits)
for cfg in ${CFGS:-128:0 96:2 64:3 128:1 256:1 192:1}; do
  b=${cfg%%:*}; c=${cfg##*:}
  echo -n "block $b ctas/SM $c: "
  B200DISORT_BLOCK=$b B200DISORT_CTAS_PER_SM=$c python bench.py --workload gcm --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms'%(d['value'], d['kernel']['avg_ms']))"
done
