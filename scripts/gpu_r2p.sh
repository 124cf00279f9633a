mkdir -p gpurun_out
python scripts/trace.py 2>&1 | awk '/iteration 2/{p=1} p' > gpurun_out/r2p_trace.txt; head -12 gpurun_out/r2p_trace.txt
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/r2p_bench.json 2>gpurun_out/r2p_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2p_bench.json')); print({k:j[k] for k in ('value','ms_per_step','value_l2_flushed','gpu_launches')}, 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['ms_per_step_host_wall'])"; tail -3 gpurun_out/r2p_bench.err
