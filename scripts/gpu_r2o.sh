mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|error|passed|failed" | head -8 > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
python scripts/trace.py 2>&1 | awk '/iteration 2/{p=1} p' > gpurun_out/r2o_trace.txt; head -16 gpurun_out/r2o_trace.txt
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/r2o_bench.json 2>gpurun_out/r2o_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2o_bench.json')); print({k:j[k] for k in ('value','ms_per_step','value_l2_flushed','ms_per_step_l2_flushed','gpu_launches')}, j['e2e']['value'], j['run']['l2'][:60], j['clocks'])"; tail -3 gpurun_out/r2o_bench.err
