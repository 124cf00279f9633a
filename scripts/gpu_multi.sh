mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/../tests/dp_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -9
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r5_bench_g${N}_peer.json 2> gpurun_out/r5_bench_g.err
python -c "import json;d=json.load(open('gpurun_out/r5_bench_g${N}_peer.json'));print('peer',d['n_gpus'],d['value'],d['ms_per_step'])"; tail -2 gpurun_out/r5_bench_g.err | grep -v OMP
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 100 --warmup 10 --dp nccl > gpurun_out/r5_bench_g${N}_nccl.json 2>> gpurun_out/r5_bench_g.err
python -c "import json;d=json.load(open('gpurun_out/r5_bench_g${N}_nccl.json'));print('nccl',d['n_gpus'],d['value'],d['ms_per_step'])"
