mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|error|passed|failed|flips" | head -8 > gpurun_out/r2q_pytest.log; cat gpurun_out/r2q_pytest.log
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/r2q_bench.json 2>gpurun_out/r2q_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2q_bench.json')); print({k:j[k] for k in ('value','ms_per_step','value_l2_flushed','gpu_launches')}, 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['ms_per_step_host_enqueue'], j['e2e']['launch'])"; tail -3 gpurun_out/r2q_bench.err
