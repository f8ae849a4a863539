#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" ) > gpurun_out/pytest_gpu7.log 2>&1
tail -3 gpurun_out/pytest_gpu7.log
rm -f gpurun_out/tune_v4.jsonl
timeout 600 python tools/tune_spmm.py --config C2 --sweep epb --out gpurun_out/tune_v4.jsonl 2>/dev/null | cut -c1-230
timeout 300 python tools/tune_spmm.py --gather-sweep --iters 3 2>/dev/null | cut -c1-200
