AMB_ATTN_VER=6 AMB_PROBE_TAG=_v7n timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | cut -c1-200 | grep "ap_\|status\|sdpa"
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|passed|failed|^FAILED|^E  " | cut -c1-260 | head -40
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; tail -c 300 gpurun_out/bench_v7.json; tail -3 gpurun_out/bench_v7.err | cut -c1-300
