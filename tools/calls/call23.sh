set -x
timeout 300 python tools/gemm_one.py 2>&1 | grep GEMM_ONE
AMB_PROBE_TAG=_g2 timeout 600 python tools/gpu_probe.py gemm gemm_perf 2>&1 | grep -E "status|p_|cublas|EXC"
