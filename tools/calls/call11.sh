set -x
AMB_PROBE_TAG=_o1 timeout 600 python tools/gpu_probe.py attn attn_more attn_perf 2>&1 | grep -v "^  a_\|^  am_" 
for V in o0 o1r; do AMB_PROBE_LIB=variants/libv_$V.so AMB_PROBE_TAG=_$V timeout 300 python tools/gpu_probe.py attn_more attn_perf 2>&1 | grep -E "status|ap_s32784|EXC"; done
AMB_PROBE_TAG=_o1b timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | grep -E "status|ap_s32784|EXC"
AMB_PROBE_LIB=variants/libv_o1t.so timeout 300 python tools/attn_trace.py 2>&1 | tail -4
