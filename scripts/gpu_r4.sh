mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
cat gpurun_out/r4_bench.json; tail -3 gpurun_out/r4_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r4_launches_dram.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4_ncu_bench.log 2>&1
