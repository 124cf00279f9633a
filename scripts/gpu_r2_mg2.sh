mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|BAD|FAIL|Error" | head -10 > gpurun_out/mg2c_pytest.log; cat gpurun_out/mg2c_pytest.log
python scripts/trace.py --dump 4 2>&1 | awk '/iteration 2/{p=1} p' > gpurun_out/mg2c_trace1.txt; grep "span" gpurun_out/mg2c_trace1.txt
timeout 300 python bench.py --steps 3000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('N=1', {k:j[k] for k in ('value','ms_per_step','value_l2_flushed')}, 'e2e', j['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('N=2', {k:j[k] for k in ('value','ms_per_step','value_l2_flushed')}, 'e2e', j['e2e']['value'], j['run']['replicas'])"
