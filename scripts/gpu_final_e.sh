mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r9_pytest.log; cat gpurun_out/r9_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; cat gpurun_out/r9_bench.json; tail -2 gpurun_out/r9_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r9_bench_reference.json 2>> gpurun_out/r9_bench.err; cut -c1-300 gpurun_out/r9_bench_reference.json
timeout 300 python bench.py --config C2A --steps 200 --warmup 20 > gpurun_out/r9_bench_c2a.json 2>> gpurun_out/r9_bench.err; python -c "import json;d=json.load(open('gpurun_out/r9_bench_c2a.json'));print('C2A',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('cpu_baseline'))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r9_launches_dram.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r9_launches_dram_c2a.csv python bench.py --config C2A --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu_bench_c2a.log 2>&1
for k in k_exchange_fwd_fast k_exchange_bwd; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r9_c2a_$k python bench.py --config C2A --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu_c2a_$k.log 2>&1
done
