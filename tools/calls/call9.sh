AMB_ATTN_VER=6 AMB_ATTN_EMU=1 AMB_PROBE_TAG=_v7p1 timeout 400 python tools/gpu_probe.py attn_more attn_perf 2>&1 | cut -c1-200 | grep "am_\|ap_\|status\|sdpa"
AMB_ATTN_VER=6 AMB_ATTN_EMU=0 AMB_PROBE_TAG=_v7p0 timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | cut -c1-200 | grep "ap_s32\|status"
AMB_ATTN_VER=6 AMB_ATTN_EMU=2 AMB_PROBE_TAG=_v7p2 timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | cut -c1-200 | grep "ap_s32\|status"
AMB_PROBE_TAG=_r3 timeout 600 python tools/gpu_probe.py gemm gemm_perf 2>&1 | cut -c1-200 | grep "g_\|p_\|status\|cublas"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED|^E  " | cut -c1-260 | head -30
timeout 900 python bench.py --steps 4 --warmup 3 --no-video --no-eager --no-cpu-baseline > gpurun_out/bench_v7p.json 2> gpurun_out/bench_v7p.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v7p.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, {k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms','share_of_step')}, d['roofline_gemm']['achieved'], d['clocks']['sm_mhz'])
PY
tail -2 gpurun_out/bench_v7p.err | cut -c1-300
