#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" ) > gpurun_out/pytest_gpu4.log 2>&1
tail -3 gpurun_out/pytest_gpu4.log
for v in "" occ0 occ2; do
  export PGCN_B200_VARIANT=$v
  rm -f gpurun_out/tune_v3_$v.jsonl gpurun_out/gather_v3_$v.jsonl
  timeout 600 python tools/tune_spmm.py --config C2 --sweep small --out gpurun_out/tune_v3_$v.jsonl > gpurun_out/tune_v3_$v.log 2>&1
  timeout 300 python tools/tune_spmm.py --gather-sweep --iters 3 --out gpurun_out/gather_v3_$v.jsonl > gpurun_out/gather_v3_$v.log 2>&1
  echo "== variant '$v'"
  python - <<PY
import json
rows=[json.loads(l) for l in open('gpurun_out/tune_v3_$v.jsonl')]
pts=sorted([r for r in rows if 'edges_per_block' in r], key=lambda r:r['ms'])
for r in pts[:6]: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in('edges_per_block','tile_floats','unroll','ms','frac','gather_GBs')})
for l in open('gpurun_out/gather_v3_$v.jsonl'):
    r=json.loads(l); print(r['window_MB'], r['tile_floats'], r['unroll'], round(r['ms'],3), round(r['gather_GBs']))
PY
done
