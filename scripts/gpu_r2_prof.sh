# round-2 evidence run (1 GPU): parity, smoke, bench (both arms), ncu launch lists (default cache control and --cache-control none),
# ncu --set full of the kernels with tensor-pipe counters, in-situ kernel times and the per-CTA timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; cut -c1-260 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 300 python bench.py --config C3 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2>> gpurun_out/r02_bench.err; cut -c1-160 gpurun_out/r02_bench_c3.json
timeout 300 python bench.py --config C2A --no-cpu-baseline > gpurun_out/r02_bench_c2a_desc_attn.json 2>> gpurun_out/r02_bench.err; cut -c1-160 gpurun_out/r02_bench_c2a_desc_attn.json
timeout 300 python scripts/ktime.py --config C2 --iters 60 > gpurun_out/r02_ktime_c2.txt 2>&1; tail -16 gpurun_out/r02_ktime_c2.txt
timeout 300 python scripts/trace.py > gpurun_out/r02_trace_c2.txt 2>&1; grep span gpurun_out/r02_trace_c2.txt | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_dram_ncu.csv python scripts/steps.py --iters 12 > /dev/null 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_dram_warm_ncu.csv python scripts/steps.py --iters 12 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_dram_ncu.csv 2>&1 | tail -9
python scripts/launch_summary.py gpurun_out/r02_launches_dram_warm_ncu.csv 2>&1 | tail -9
for k in k_pre k_wgrad k_exchange_fwd_fast k_exchange_bwd_fast k_baseline_fwd; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/r02_$k python scripts/steps.py --iters 5 > /dev/null 2>&1
ncu -i /tmp/r02_$k.ncu-rep --page details > gpurun_out/r02_${k}_ncu_details.txt 2>&1
ncu -i /tmp/r02_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; val=rows[-1]
keep=[i for i,h in enumerate(hdr) if any(s in h for s in ('pipe_tensor','dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__warps_active.avg.pct','launch__registers_per_thread','sm__inst_executed_pipe_tensor','smsp__inst_executed.sum','dram__throughput.avg.pct','lts__t_bytes.sum'))]
for i in keep: print(hdr[i], '=', val[i], rows[1][i] if len(rows)>2 else '')
" > gpurun_out/r02_${k}_ncu_raw_excerpt.txt
ncu -i /tmp/r02_$k.ncu-rep --page source --print-source sass --csv > /tmp/r02_$k.sass.csv 2>/dev/null
python scripts/ncu_stalls.py /tmp/r02_$k.sass.csv 12 > gpurun_out/r02_${k}_stalls.txt 2>&1
done
grep -i tensor gpurun_out/r02_k_pre_ncu_raw_excerpt.txt | head
du -sh gpurun_out
