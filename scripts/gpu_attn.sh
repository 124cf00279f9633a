mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; python -c "import json;d=json.load(open('gpurun_out/r8_bench.json'));print(d['value'],d['ms_per_step'],d['ms_per_step_l2_warm'],d['e2e']['value'])"
