python scripts/phase_timing_bwd.py 2>&1 | tail -10
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
cut -c1-330 gpurun_out/r3_bench.json; tail -2 gpurun_out/r3_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1
