mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "data_parallel" 2>&1 | grep -E "passed|failed|BAD|FAIL" | head -20 > gpurun_out/mg8_pytest.log; cat gpurun_out/mg8_pytest.log
for cfg in C4 C5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --config $cfg --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/mg8_bench_$cfg.json 2> gpurun_out/mg8_bench_$cfg.err; python -c "
import json; j=json.load(open('gpurun_out/mg8_bench_$cfg.json')); print('$cfg', {k:j[k] for k in ('value','ms_per_step','value_l2_flushed','n_gpus')}, 'e2e', j['e2e']['value'], j['run']['replicas'], j['run']['dp_oracle_check'])"; grep -v "^\[W\|^$\|Warning\|warn\|OMP\|\*\*\*" gpurun_out/mg8_bench_$cfg.err | tail -4
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/mg8_bench_g4.json 2> gpurun_out/mg8_bench_g4.err; python -c "
import json; j=json.load(open('gpurun_out/mg8_bench_g4.json')); print('N=4', {k:j[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', j['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 scripts/trace.py --config C4 2>&1 | grep -v "^\[W\|^$\|Warning\|warn\|\*\*\*\|OMP_NUM" | awk "/iteration 2/{p=1} p" > gpurun_out/mg8_trace.txt; grep "span" gpurun_out/mg8_trace.txt
