mkdir -p gpurun_out
for k in k_wgrad k_pre k_exchange_bwd_fast k_exchange_fwd_fast; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r3_$k python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_$k.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1
