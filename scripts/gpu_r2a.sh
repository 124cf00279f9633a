mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
timeout 300 python bench.py > gpurun_out/r2a_bench_default.json 2>> gpurun_out/r2a_bench.err; cut -c1-400 gpurun_out/r2a_bench_default.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_reference.json 2>> gpurun_out/r2a_bench.err; cut -c1-300 gpurun_out/r2a_bench_reference.json
