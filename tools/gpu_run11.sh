#!/bin/bash
mkdir -p gpurun_out
for v in "" ld0 ""; do
export PGCN_B200_VARIANT=$v
echo "== variant '$v'"
timeout 600 python tools/tune_spmm.py --config C2 --sweep depth --iters 10 2>/dev/null | grep edges_per_block | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l); print(r['edges_per_block'], round(r['ms'],4), round(r['frac'],4))
"
done
