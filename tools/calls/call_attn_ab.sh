set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
export AMB_WAIT_TIMEOUT=1
AMB_ATTN_VER=6 AMB_ATTN_EMU=2 AMB_PROBE_TAG=_v6e2 timeout 900 python tools/gpu_probe.py attn attn_more attn_perf
AMB_ATTN_VER=6 AMB_ATTN_EMU=1 AMB_PROBE_TAG=_v6e1 timeout 300 python tools/gpu_probe.py attn_perf
AMB_ATTN_VER=6 AMB_ATTN_EMU=3 AMB_PROBE_TAG=_v6e3 timeout 300 python tools/gpu_probe.py attn_perf
AMB_ATTN_VER=6 AMB_ATTN_EMU=0 AMB_PROBE_TAG=_v6e0 timeout 300 python tools/gpu_probe.py attn_perf
AMB_ATTN_VER=4 AMB_PROBE_TAG=_v4 timeout 300 python tools/gpu_probe.py attn_more attn_perf
