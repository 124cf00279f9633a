mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_exchange_fwd_fast -s 3 -c 1 -f -o gpurun_out/r2_fwd_fast python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_exchange_bwd_fast -s 3 -c 1 -f -o gpurun_out/r2_bwd_fast python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad -s 3 -c 1 -f -o gpurun_out/r2_wgrad python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_baseline_fwd -s 3 -c 1 -f -o gpurun_out/r2_bas python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bas.log 2>&1
