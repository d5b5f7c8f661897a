set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'flash_attn|gemm|layernorm|cfg_euler|cast_|add_bias|timestep|attn_small|split3|patchify|rope' -c 1400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-video > gpurun_out/r02_launches_bench.log 2>&1; tail -c 200 gpurun_out/r02_launches_bench.log
