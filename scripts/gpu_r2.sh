mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
cat gpurun_out/r2_pytest.log gpurun_out/r2_bench.json
tail -3 gpurun_out/r2_bench.err
grep -E "k_" gpurun_out/r2_launches.csv | tail -12 | awk -F'","' '{print $5, $NF}'
