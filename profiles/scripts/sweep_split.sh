# e36: layer-split form against the fused kernel at nstr = 16 (GCM shape, full size), adding phase as threads / teams, chunk sizes
run() { python bench.py --workload gcm --steps 3 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms clocks %s'%(d['value'], d['kernel']['avg_ms'], d['clocks']))"; }
echo -n "fused                   : "; B200DISORT_SPLIT=0 run
echo -n "split, thread adding    : "; B200DISORT_SPLIT=1 B200DISORT_TEAM=0 run
echo -n "split, team adding      : "; B200DISORT_SPLIT=1 run
echo -n "split, team (1 wave)    : "; B200DISORT_SPLIT=1 B200DISORT_SPLIT_CHUNK=37888 run
echo -n "split, team (4 waves)   : "; B200DISORT_SPLIT=1 B200DISORT_SPLIT_CHUNK=151552 run
