mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3000 --warmup 20 --no-cpu-baseline > gpurun_out/mg8b_bench_C4.json 2> gpurun_out/mg8b_bench_C4.err; python -c "
import json; j=json.load(open('gpurun_out/mg8b_bench_C4.json')); print('C4', {k:j[k] for k in ('value','ms_per_step','value_l2_flushed','n_gpus')}, 'e2e', j['e2e']['value'], j['run']['replicas'])"; grep -v "^\[W\|^$\|Warning\|warn\|OMP\|\*\*\*" gpurun_out/mg8b_bench_C4.err | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 scripts/trace.py --config C4 2>&1 | grep -v "^\[W\|^$\|Warning\|warn\|\*\*\*\|OMP_NUM" | awk "/iteration 2/{p=1} p" > gpurun_out/mg8b_trace.txt; grep "span" gpurun_out/mg8b_trace.txt; grep -A7 "k_update\|k_peer" gpurun_out/mg8b_trace.txt | head -18
timeout 300 python bench.py --steps 3000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('N=1', {k:j[k] for k in ('value','ms_per_step','value_l2_flushed')}, 'e2e', j['e2e']['value'])"
