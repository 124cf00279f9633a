mkdir -p gpurun_out
for k in k_exchange_fwd_fast; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r8_$k python bench.py --config C2A --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_ncu_$k.log 2>&1
done
