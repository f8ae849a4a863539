#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" ) > gpurun_out/pytest_gpu5.log 2>&1
tail -3 gpurun_out/pytest_gpu5.log
rm -f gpurun_out/tune_v4.jsonl
for cfg in "ld0:64" "ld1:64" ":32" ":64" ":96" "ld3:64" "ld3:96"; do
  v=${cfg%%:*}; hot=${cfg#*:}
  export PGCN_B200_VARIANT=$v PGCN_HOT_MB=$hot
  echo "== variant '$v' hot_mb=$hot"
  timeout 600 python tools/tune_spmm.py --config C2 --sweep mini 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l)
    if 'edges_per_block' in r: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in('edges_per_block','unroll','ms','frac','gather_GBs')})
"
done
