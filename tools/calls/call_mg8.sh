# 8-GPU call: bench (sharded window is the scored value) with both exchanges, then c5 and c4 at their named shapes
N=8
TR="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for X in nccl peer; do
  $TR bench.py --gpus $N --steps 4 --warmup 3 --exchange $X > gpurun_out/bench_${N}gpu_$X.json 2> gpurun_out/bench_${N}gpu_$X.err
  tail -c 1200 gpurun_out/bench_${N}gpu_$X.json; echo; grep -v "^\[W\|^W0\|Warning\|warn" gpurun_out/bench_${N}gpu_$X.err | tail -5 | cut -c1-300
done
for X in nccl peer; do
  $TR tools/run_config.py c5 --exchange $X > gpurun_out/c5_$X.json 2> gpurun_out/c5_$X.err
  tail -c 800 gpurun_out/c5_$X.json; echo; grep -v "^\[W\|^W0\|Warning\|warn" gpurun_out/c5_$X.err | tail -5 | cut -c1-300
done
$TR tools/run_config.py c4 --clips 16 > gpurun_out/c4.json 2> gpurun_out/c4.err
tail -c 800 gpurun_out/c4.json; echo; grep -v "^\[W\|^W0\|Warning\|warn" gpurun_out/c4.err | tail -5 | cut -c1-300
