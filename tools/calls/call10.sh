set -x
S=32784 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pair -s 2 -c 1 -f -o gpurun_out/ncu_pair_v7 python tools/attn_one.py > gpurun_out/ncu_pair_v7.log 2>&1; tail -2 gpurun_out/ncu_pair_v7.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-video > gpurun_out/r02_launches_bench.log 2>&1; tail -c 300 gpurun_out/r02_launches_bench.log
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "SHARD_EMU|CHAMFER|DEFAULT_CONFIG|passed|failed|FAILED|Error" | cut -c1-400
