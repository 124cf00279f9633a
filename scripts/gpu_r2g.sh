mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3_desc_attn" 2>&1 | grep -E "Error|error|passed|failed" | head -8 > gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_pytest.log
for k in k_pre k_baseline_fwd k_wgrad k_update; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/r2g_$k python scripts/steps.py --iters 5 > gpurun_out/r2g_ncu_$k.log 2>&1
ncu -i /tmp/r2g_$k.ncu-rep --page details > gpurun_out/r2g_${k}_details.txt 2>&1
ncu -i /tmp/r2g_$k.ncu-rep --page source --print-source cuda,sass --csv > /tmp/r2g_$k.src.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/r2g_$k.src.csv 45 > gpurun_out/r2g_${k}_hot_lines.txt 2>&1
done
ls -la gpurun_out | tail -12
