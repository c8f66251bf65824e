# e42: what halving the back-sweep record traffic could buy at nstr = 8 (experiment build: wrong fluxes, same instructions)
for v in "" _halfrec; do
  lib=$PWD/pyharp_b200/libb200disort$v.so
  for w in amars_lw amars_sw; do
    echo -n "$(basename $lib) $w: "
    B200DISORT_LIB=$lib python bench.py --workload $w --prop-level --steps 10 --warmup 3 --no-cpu --no-e2e --no-others 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g solves/s kernel %.3f ms'%(d['value'], d['kernel']['avg_ms']))"
  done
done
lib=$PWD/pyharp_b200/libb200disort_halfrec.so
B200DISORT_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:disort_flux_kernel -s 3 -c 1 python bench.py --workload amars_lw --prop-level --steps 2 --warmup 3 --no-cpu --no-e2e --no-others 2>&1 | grep -E "dram__|gpu__time"
