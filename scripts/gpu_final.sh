mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r6_pytest.log; cat gpurun_out/r6_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; cat gpurun_out/r6_bench.json; tail -2 gpurun_out/r6_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r6_bench_reference.json 2>> gpurun_out/r6_bench.err; cut -c1-300 gpurun_out/r6_bench_reference.json
timeout 300 python bench.py --config C3 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r6_bench_c3.json 2>> gpurun_out/r6_bench.err; python -c "import json;d=json.load(open('gpurun_out/r6_bench_c3.json'));print('C3',d['value'],d['ms_per_step'],d['config']['exchange_steps'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r6_launches_dram.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu_bench.log 2>&1
for k in k_exchange_fwd_fast k_wgrad k_exchange_bwd_fast; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r6_$k python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu_$k.log 2>&1
done
python scripts/phase_timing.py libmmg_dbg.so 2>&1 | tail -9 > gpurun_out/r6_phase_timing_fwd.txt; python scripts/phase_timing_bwd.py 2>&1 | tail -10 > gpurun_out/r6_phase_timing_bwd.txt; cat gpurun_out/r6_phase_timing_fwd.txt
