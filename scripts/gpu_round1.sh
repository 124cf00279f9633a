mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r1_bench_ref.json 2>> gpurun_out/r1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_exchange -s 6 -c 2 -o gpurun_out/r1_prof_exchange python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
cat gpurun_out/r1_pytest.log gpurun_out/r1_bench.json gpurun_out/r1_bench_ref.json
tail -3 gpurun_out/r1_bench.err
