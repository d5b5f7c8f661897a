N=4
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}gpu_final.json 2> gpurun_out/bench_${N}gpu_final.err
tail -c 700 gpurun_out/bench_${N}gpu_final.json; echo; grep -v "^\[W\|^W0\|Warning\|warn" gpurun_out/bench_${N}gpu_final.err | tail -5 | cut -c1-300
