AMB_ATTN_VER=6 AMB_ATTN_EMU=1 AMB_PROBE_TAG=_v7p1 timeout 400 python tools/gpu_probe.py attn_more attn_perf 2>&1 | cut -c1-200 | grep "am_\|ap_\|status\|sdpa"
AMB_ATTN_VER=6 AMB_ATTN_EMU=0 AMB_PROBE_TAG=_v7p0 timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | cut -c1-200 | grep "ap_s32\|status"
AMB_ATTN_VER=6 AMB_ATTN_EMU=2 AMB_PROBE_TAG=_v7p2 timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | cut -c1-200 | grep "ap_s32\|status"
