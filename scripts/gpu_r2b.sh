mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cut -c1-600 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2b_bench_g2.json 2> gpurun_out/r2b_bench_g2.err; cut -c1-2500 gpurun_out/r2b_bench_g2.json; tail -5 gpurun_out/r2b_bench_g2.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2b_launches.csv 2>&1 | tail -20
