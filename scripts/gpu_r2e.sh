mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_pytest.log
timeout 300 python scripts/ktime.py --config C2 2>&1 | tee gpurun_out/r2e_ktime_c2.txt
MMG_UMMA=0 timeout 300 python scripts/ktime.py --config C2 2>&1 | tee gpurun_out/r2e_ktime_c2_noumma.txt
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; cut -c1-300 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
