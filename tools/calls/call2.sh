set -x
export AMB_ATTN_VER=6 AMB_ATTN_EMU=1
timeout 120 python tools/attn_trace.py > gpurun_out/trace_v6.log 2>&1; tail -20 gpurun_out/trace_v6.log
S=16392 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pair -s 1 -c 1 -f -o gpurun_out/ncu_pair python tools/attn_one.py > gpurun_out/ncu_pair.log 2>&1; tail -3 gpurun_out/ncu_pair.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/gputests_call2.log; cat gpurun_out/gputests_call2.log
