mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== fused"; timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r5_bench_fused.json 2> gpurun_out/r5_bench.err; python -c "import json;d=json.load(open('gpurun_out/r5_bench_fused.json'));print(d['value'],d['ms_per_step'],d['ms_per_step_l2_warm'],d['e2e']['value'],d['gpu_launches'])"
echo "== unfused"; MMG_FUSED_UPDATE=0 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r5_bench_unfused.json 2>> gpurun_out/r5_bench.err; python -c "import json;d=json.load(open('gpurun_out/r5_bench_unfused.json'));print(d['value'],d['ms_per_step'],d['ms_per_step_l2_warm'],d['e2e']['value'])"
tail -3 gpurun_out/r5_bench.err
