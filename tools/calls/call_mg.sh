# usage: call_mg.sh N   (multi-GPU: sharded-window parity over NCCL + peer copies, then the bench with both exchanges)
N=$1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_window_shard_gpu.py -m gpu -q -k "matches_single_gpu" 2>&1 | tail -5 | cut -c1-400
fi
for X in nccl peer; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --exchange $X > gpurun_out/bench_${N}gpu_$X.json 2> gpurun_out/bench_${N}gpu_$X.err
  tail -c 900 gpurun_out/bench_${N}gpu_$X.json; echo; grep -v "^\[W\|^W0\|Warning\|warn" gpurun_out/bench_${N}gpu_$X.err | tail -5 | cut -c1-300
done
