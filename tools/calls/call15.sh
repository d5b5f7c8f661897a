set -x
AMB_PROBE_TAG=_m3 timeout 300 python tools/gpu_probe.py attn attn_more attn_perf 2>&1 | grep -v "^  a_\|^  am_" 
for V in head; do AMB_PROBE_LIB=variants/libv_$V.so AMB_PROBE_TAG=_$V timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | grep -E "status|ap_s32784|EXC"; done
AMB_PROBE_TAG=_m3b timeout 300 python tools/gpu_probe.py attn_perf 2>&1 | grep -E "status|ap_s|sdpa|EXC"
AMB_PROBE_LIB=variants/libv_m3t.so timeout 300 python tools/attn_trace.py 2>&1 | tail -4
