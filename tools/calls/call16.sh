set -x
S=32784 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pair -s 2 -c 1 -f -o gpurun_out/ncu_pair_v8 python tools/attn_one.py > gpurun_out/ncu_pair_v8.log 2>&1; tail -2 gpurun_out/ncu_pair_v8.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; tail -c 2500 gpurun_out/bench_v8.json; tail -3 gpurun_out/bench_v8.err
