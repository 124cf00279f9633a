mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/probe_symm.py 2>&1 | grep -i "multicast\|handle attrs" | head -4
echo "(DP pytest skipped in this run)"
for nv in 1 0; do
MMG_NVLS=$nv timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 3000 --warmup 20 --no-cpu-baseline 2>gpurun_out/nvls.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('NVLS=$nv N=$N', {k:j[k] for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], j['run']['replicas'], j['run']['dp_oracle_check']['param_max_err_over_lr'])"; grep -v "^\[W\|^$\|Warning\|warn\|OMP\|\*\*\*\|NCCL version" gpurun_out/nvls.err | tail -3
done
MMG_NVLS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 scripts/trace.py --config C4 2>&1 | grep -v "^\[W\|^$\|Warning\|warn\|\*\*\*\|OMP_NUM" | awk "/iteration 2/{p=1} p" > gpurun_out/nvls_trace.txt; grep -A6 "k_peer" gpurun_out/nvls_trace.txt | head -8; grep "k_update.*span" gpurun_out/nvls_trace.txt
