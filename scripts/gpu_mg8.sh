mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r9_bench_g8.json 2> gpurun_out/r9_bench_g8.err
python -c "import json;d=json.load(open('gpurun_out/r9_bench_g8.json'));print('C4 x8',d['n_gpus'],d['value'],d['ms_per_step'],d['config']['dp_reduction'][:40])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --config C2A --steps 200 --warmup 20 > gpurun_out/r9_bench_g8_c2a.json 2>> gpurun_out/r9_bench_g8.err
python -c "import json;d=json.load(open('gpurun_out/r9_bench_g8_c2a.json'));print('C2A x8',d['n_gpus'],d['value'],d['ms_per_step'])"
