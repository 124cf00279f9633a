mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; python -c "import json;d=json.load(open('gpurun_out/r5_bench.json'));print(d['value'],d['ms_per_step'],d['ms_per_step_l2_warm'],d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu_bench.log 2>&1
