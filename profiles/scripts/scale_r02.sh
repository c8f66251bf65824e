# Multi-GPU lines of round 2 (run on an 8-GPU box): weak scaling of the default band-level step (device value and
# end-to-end with host buffers) and strong scaling of the line-by-line spectral-point sharding with its NCCL all-reduce.
out=gpurun_out
mkdir -p $out
port=29600
for n in 8 4 2; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 20 --warmup 3 --no-others > $out/bench_lw_${n}gpu_v1.json 2> $out/bench_lw_${n}gpu_v1.err; echo "default N=$n rc=$?"
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 50 --warmup 5 --workload lbl --shard wave > $out/bench_lblwave_${n}gpu_v1.json 2> $out/bench_lblwave_${n}gpu_v1.err; echo "lbl wave N=$n rc=$?"
done
timeout 600 python -m pytest tests/test_gpu_nccl.py -q > $out/pytest_nccl_v1.log 2>&1; tail -2 $out/pytest_nccl_v1.log
