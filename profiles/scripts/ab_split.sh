# A/B of library variants on the GCM shape (layer-split form): bench value + per-kernel times from an ncu launch list
# Usage: bash profiles/scripts/ab_split.sh <lib suffix> ...   ("" = the product library)
export B200DISORT_SPLIT=1
for v in "$@"; do
  lib=$PWD/pyharp_b200/libb200disort${v:+_$v}.so
  [ "$v" = "-" ] && lib=$PWD/pyharp_b200/libb200disort.so
  echo "== $lib"
  B200DISORT_LIB=$lib python bench.py --workload gcm --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.4g solves/s kernel %.3f ms'%(d['value'], d['kernel']['avg_ms']))"
  B200DISORT_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"split_|team_" -c 45 --csv --log-file gpurun_out/ab_launches.csv python bench.py --workload gcm --ncol 16384 --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
  python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/ab_launches.csv')) if len(r)>5 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows: d[r[4].split('(')[0][-28:]].append(float(r[-1].replace(',','')))
print('  '+'; '.join('%s %.3f ms'%(k,sum(v)/len(v)/1e6) for k,v in d.items()))
PY
done
