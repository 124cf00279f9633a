mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|error|passed|failed" | head -8 > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
for x in 0 64; do
echo "=== MMG_XSKIP=$x"
MMG_XSKIP=$x timeout 300 python scripts/ktime.py --config C2 --iters 40 2>&1 | grep -A8 "flushed\|back to back"
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | cut -c1-200
