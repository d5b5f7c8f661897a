set -x
AMB_PROBE_TAG=_g1 timeout 600 python tools/gpu_probe.py gemm gemm_perf 2>&1 | grep -E "status|p_|cublas|EXC"
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|CHAMFER|DEFAULT_CONFIG|passed|failed|FAILED|Error" | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
