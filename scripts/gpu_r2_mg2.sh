mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "data_parallel" 2>&1 | grep -E "Error|error|passed|failed|PASS|FAIL|dmax|skipped" | head -20 > gpurun_out/mg2_pytest.log; cat gpurun_out/mg2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/mg2_bench.json 2> gpurun_out/mg2_bench.err; python -c "
import json; j=json.load(open('gpurun_out/mg2_bench.json')); print({k:j[k] for k in ('value','ms_per_step','value_l2_flushed','gpu_launches','n_gpus')}, 'e2e', j['e2e']['value'], j['run']['dp_reduction'][:40], j['run']['replicas'], j['run']['dp_oracle_check'])"; grep -v "^\[W\|^$\|Warning\|warn" gpurun_out/mg2_bench.err | tail -8
