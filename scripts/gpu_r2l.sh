mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|error|passed|failed" | head -8 > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
python scripts/trace.py --dump 2 2>&1 | awk '/iteration 2/{p=1} p' > gpurun_out/r2l_trace.txt; grep -A12 "k_baseline_fwd" gpurun_out/r2l_trace.txt | head -14; grep "cta" gpurun_out/r2l_trace.txt | awk 'NR%8==1' | head -30
