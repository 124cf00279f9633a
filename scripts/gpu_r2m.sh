mkdir -p gpurun_out
for k in k_baseline_fwd k_wgrad; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/r2m_$k python scripts/steps.py --iters 5 > gpurun_out/r2m_ncu_$k.log 2>&1
ncu -i /tmp/r2m_$k.ncu-rep --page source --print-source sass --csv > /tmp/r2m_$k.sass.csv 2>/dev/null
python scripts/ncu_stalls.py /tmp/r2m_$k.sass.csv 25 > gpurun_out/r2m_${k}_stalls.txt 2>&1
ncu -i /tmp/r2m_$k.ncu-rep --page details 2>/dev/null | grep -E "Stall|stall|Warp Cycles Per|Instruction|instruction" | head -30 > gpurun_out/r2m_${k}_warpstate.txt
done
head -30 gpurun_out/r2m_k_baseline_fwd_stalls.txt; cat gpurun_out/r2m_k_baseline_fwd_warpstate.txt
