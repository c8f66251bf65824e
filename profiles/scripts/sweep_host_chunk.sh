# end-to-end (host buffers) rate against the chunk size of the pipelined host path (fraction of a resident wave)
for w in amars_lw amars_sw lbl; do
for d in 1 2 4 8; do
  echo -n "$w div $d: "
  B200DISORT_HOST_CHUNK_DIV=$d python bench.py --workload $w --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e %.4g solves/s (device %.4g)'%(d['e2e']['value'], d['value']))"
done
done
