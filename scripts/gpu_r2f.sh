mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_pytest.log
for k in k_pre k_baseline_fwd k_wgrad k_update k_exchange_fwd_fast k_exchange_bwd_fast; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2f_$k python scripts/steps.py --iters 5 > gpurun_out/r2f_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
