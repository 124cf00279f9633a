mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_exchange_fwd_fast -s 3 -c 1 -f -o gpurun_out/r3_fwd_fast python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_fwd.log 2>&1
