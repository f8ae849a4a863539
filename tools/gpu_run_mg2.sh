#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( timeout 900 python -m pytest tests/test_multigpu.py -x -q -k "two_gpus" ) > gpurun_out/pytest_mg2.log 2>&1
tail -15 gpurun_out/pytest_mg2.log
for tr in nccl p2p; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
     bench.py --gpus 2 --steps 20 --warmup 5 --transport $tr --partition rp > gpurun_out/bench_n2_$tr.json 2> gpurun_out/bench_n2_$tr.err
  cat gpurun_out/bench_n2_$tr.json | cut -c1-1200
  tail -3 gpurun_out/bench_n2_$tr.err
done
