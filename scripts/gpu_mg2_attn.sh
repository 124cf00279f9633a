mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/../tests/dp_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -11
for mode in peer; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config C2A --steps 200 --warmup 20 --dp $mode > gpurun_out/r9_bench_g2_c2a_$mode.json 2> gpurun_out/r9_bench_g2.err
python -c "import json;d=json.load(open('gpurun_out/r9_bench_g2_c2a_$mode.json'));print('C2A $mode',d['n_gpus'],d['value'],d['ms_per_step'],d['ms_per_step_l2_warm'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 200 --warmup 20 --dp $mode > gpurun_out/r9_bench_g2_$mode.json 2>> gpurun_out/r9_bench_g2.err
python -c "import json;d=json.load(open('gpurun_out/r9_bench_g2_$mode.json'));print('C4 $mode',d['n_gpus'],d['value'],d['ms_per_step'],d['ms_per_step_l2_warm'])"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k two_gpus 2>&1 | tail -2
