export AMB_ATTN_VER=4
AMB_PROBE_TAG=_r2 timeout 900 python tools/gpu_probe.py gemm gemm_perf 2>&1 | cut -c1-230 | head -60
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|CHAMFER|DEFAULT_CONFIG|passed|failed|^FAILED|^E  " | cut -c1-300 | head -60
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; tail -c 2500 gpurun_out/bench_v4.json; tail -3 gpurun_out/bench_v4.err | cut -c1-300
