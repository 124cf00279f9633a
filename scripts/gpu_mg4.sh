mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 scripts/../tests/dp_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 200 --warmup 20 > gpurun_out/r9_bench_g4.json 2> gpurun_out/r9_bench_g4.err
python -c "import json;d=json.load(open('gpurun_out/r9_bench_g4.json'));print('C4 x4',d['n_gpus'],d['value'],d['ms_per_step'],d['config']['dp_reduction'][:40])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 4 --impl reference --steps 5 --warmup 1 2>/dev/null | cut -c1-200
