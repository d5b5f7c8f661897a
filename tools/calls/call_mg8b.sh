# 8-GPU call with the final build: bench (sharded window = value; peer exchange default), then c5
N=8
TR="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}gpu_final.json 2> gpurun_out/bench_${N}gpu_final.err
tail -c 1500 gpurun_out/bench_${N}gpu_final.json; echo; grep -v "^\[W\|^W0\|Warning\|warn" gpurun_out/bench_${N}gpu_final.err | tail -5 | cut -c1-300
$TR tools/run_config.py c5 > gpurun_out/c5_final.json 2> gpurun_out/c5_final.err
tail -c 800 gpurun_out/c5_final.json; echo; grep -v "^\[W\|^W0\|Warning\|warn" gpurun_out/c5_final.err | tail -5 | cut -c1-300
