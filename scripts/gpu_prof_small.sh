mkdir -p gpurun_out
for k in k_update k_reduce_norm k_lossgrad k_stats k_baseline_fwd; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r7_$k python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r7_ncu_$k.log 2>&1
done
