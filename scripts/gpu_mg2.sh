mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/check_peer_dp.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r4_bench_g2_peer.json 2> gpurun_out/r4_bench_g2.err
cut -c1-250 gpurun_out/r4_bench_g2_peer.json; tail -3 gpurun_out/r4_bench_g2.err | grep -v OMP
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 100 --warmup 10 --dp nccl > gpurun_out/r4_bench_g2_nccl.json 2>> gpurun_out/r4_bench_g2.err
cut -c1-250 gpurun_out/r4_bench_g2_nccl.json
