set -x
export AMB_ATTN_VER=6
AMB_ATTN_EMU=1 AMB_PROBE_TAG=_v6b1 timeout 600 python tools/gpu_probe.py attn attn_more attn_perf
AMB_ATTN_EMU=2 AMB_PROBE_TAG=_v6b2 timeout 300 python tools/gpu_probe.py attn_perf
AMB_ATTN_EMU=0 AMB_PROBE_TAG=_v6b0 timeout 300 python tools/gpu_probe.py attn_perf
AMB_ATTN_EMU=1 timeout 120 python tools/attn_trace.py > gpurun_out/trace_v6b.log 2>&1; tail -12 gpurun_out/trace_v6b.log
AMB_ATTN_EMU=1 timeout 600 python -m pytest tests/test_window_shard_gpu.py -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|passed|failed|Error" 
AMB_ATTN_VER=4 timeout 600 python -m pytest tests/test_window_shard_gpu.py -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|passed|failed|Error"
AMB_ATTN_EMU=1 timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -15
AMB_PROBE_TAG=_r2 timeout 900 python tools/gpu_probe.py gemm gemm_perf
