mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r3_bench_g2.json 2> gpurun_out/r3_bench_g2.err
cut -c1-400 gpurun_out/r3_bench_g2.json; tail -5 gpurun_out/r3_bench_g2.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r3_bench_g1.json 2>> gpurun_out/r3_bench_g2.err
cut -c1-200 gpurun_out/r3_bench_g1.json
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
