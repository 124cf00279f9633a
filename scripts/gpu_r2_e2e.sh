mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_pipeline or fused_train_step or golden_train" 2>&1 | grep -E "passed|failed|Error" | head
timeout 300 python bench.py --steps 3000 --warmup 20 --no-cpu-baseline 2>gpurun_out/e2e.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('N=1', {k:j[k] for k in ('value','ms_per_step','value_l2_flushed')}, 'e2e', j['e2e'])"; tail -3 gpurun_out/e2e.err
