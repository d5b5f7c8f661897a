set -x
timeout 300 python tools/gemm_one.py 2>&1 | grep GEMM_ONE
ONLY=fp32_res_alias_out2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 3 -c 1 -f -o gpurun_out/ncu_gemm2_res python tools/gemm_one.py > gpurun_out/ncu_gemm2_res.log 2>&1; tail -2 gpurun_out/ncu_gemm2_res.log
