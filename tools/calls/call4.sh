set -x
export AMB_ATTN_VER=6 AMB_ATTN_EMU=1
S=16392 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pair -s 1 -c 1 -f -o gpurun_out/ncu_pair_v6b python tools/attn_one.py > gpurun_out/ncu_pair_v6b.log 2>&1; tail -2 gpurun_out/ncu_pair_v6b.log
AMB_PROBE_TAG=_r2 timeout 900 python tools/gpu_probe.py gemm gemm_perf
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|CHAMFER|DEFAULT_CONFIG|passed|failed|FAILED|Error" | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-eager > gpurun_out/bench_v6b.json 2> gpurun_out/bench_v6b.err; tail -c 1500 gpurun_out/bench_v6b.json; tail -3 gpurun_out/bench_v6b.err
AMB_ATTN_VER=4 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-eager --no-video > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; tail -c 1200 gpurun_out/bench_v4.json; tail -3 gpurun_out/bench_v4.err
