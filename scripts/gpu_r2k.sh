mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|error|passed|failed" | head -8 > gpurun_out/r2k_pytest.log; cat gpurun_out/r2k_pytest.log
python scripts/trace.py 2>&1 | awk '/iteration 2/{p=1} p' > gpurun_out/r2k_trace.txt; head -70 gpurun_out/r2k_trace.txt
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | cut -c1-200
