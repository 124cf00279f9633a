mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "attn" 2>&1 | tail -3
timeout 300 python bench.py --config C2A --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r8_bench_c2a.json 2> gpurun_out/r8_bench_c2a.err; python -c "import json;d=json.load(open('gpurun_out/r8_bench_c2a.json'));print('C2A',d['value'],d['ms_per_step'],d['ms_per_step_l2_warm'],d['e2e']['value'])"; tail -3 gpurun_out/r8_bench_c2a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r8_launches_c2a.csv python bench.py --config C2A --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r8_ncu_c2a.log 2>&1
MMG_FORCE_GENERIC=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r8_launches_c2gen.csv python bench.py --config C2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r8_ncu_c2gen.log 2>&1
