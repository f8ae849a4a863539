#!/bin/bash
# first GPU pass: parity tests, option sweep, L2 probe, bench line, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/tune_spmm.py --config C2 --sweep default --baselines --out gpurun_out/tune_c2.jsonl > gpurun_out/tune_c2.log 2>&1
tail -3 gpurun_out/tune_c2.log
timeout 300 python tools/tune_spmm.py --gather-sweep --iters 5 --out gpurun_out/gather.jsonl > gpurun_out/gather.log 2>&1
tail -3 gpurun_out/gather.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
