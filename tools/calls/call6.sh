export AMB_ATTN_VER=6
AMB_ATTN_EMU=1 AMB_PROBE_TAG=_v7e1 timeout 600 python tools/gpu_probe.py attn attn_more attn_perf 2>&1 | cut -c1-200 | grep -v "^$" | head -70
AMB_ATTN_EMU=2 AMB_PROBE_TAG=_v7e2 timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | cut -c1-200 | grep "ap_\|status"
AMB_ATTN_EMU=0 AMB_PROBE_TAG=_v7e0 timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | cut -c1-200 | grep "ap_\|status"
AMB_ATTN_MODE=0 AMB_PROBE_TAG=_v7x timeout 300 python tools/gpu_probe.py attn_more 2>&1 | cut -c1-200 | grep "am_\|status"
AMB_ATTN_EMU=1 timeout 120 python tools/attn_trace.py > gpurun_out/trace_v7.log 2>&1; tail -12 gpurun_out/trace_v7.log | cut -c1-250
