#!/bin/bash
mkdir -p gpurun_out
export PGCN_B200_VARIANT=ld2
timeout 900 python tools/tune_spmm.py --config C2 --hot-sweep 8,16,32,40,48,56,64,80,96,128,600 --iters 10 2>/dev/null | python -c "
import sys, json
last=None
for l in sys.stdin:
    r=json.loads(l)
    if 'ms' in r: last=r
    if 'hot_mb' in r and last: print(r['hot_mb'], last['edges_per_block'], round(last['ms'],4))
"
