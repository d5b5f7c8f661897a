set -x
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|CHAMFER|DEFAULT_CONFIG|DINO|passed|failed|FAILED|Error" | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-video > gpurun_out/r02_launches_bench.log 2>&1; tail -c 200 gpurun_out/r02_launches_bench.log
timeout 900 python bench.py --no-cpu-baseline --no-eager > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 700 gpurun_out/bench_final2.json; tail -3 gpurun_out/bench_final2.err
