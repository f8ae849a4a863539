#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" ) > gpurun_out/pytest_gpu10.log 2>&1
tail -2 gpurun_out/pytest_gpu10.log
timeout 600 python tools/tune_spmm.py --config C2 --sweep depth --iters 10 2>/dev/null | grep depth | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l); print(r['depth'], r['edges_per_block'], round(r['ms'],4), round(r['frac'],4))
"
timeout 600 python tools/tune_spmm.py --config C2 --transpose --sweep depth --iters 10 2>/dev/null | grep depth | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l); print('T', r['depth'], r['edges_per_block'], round(r['ms'],4))
"
