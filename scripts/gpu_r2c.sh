mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cut -c1-300 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2c_bench_g2.json 2> gpurun_out/r2c_bench_g2.err; cut -c1-1800 gpurun_out/r2c_bench_g2.json; grep -v "^\[W\|^$" gpurun_out/r2c_bench_g2.err | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2c_launches.csv 2>&1 | tail -20
for k in k_wgrad k_update; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2c_$k python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
