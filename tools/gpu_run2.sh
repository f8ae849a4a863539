#!/bin/bash
set -x
mkdir -p gpurun_out
# one full capture of the top kernel (cold + warm launch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rowblock -s 3 -c 2 -f -o gpurun_out/prof_c2_r1 \
    python tools/tune_spmm.py --config C2 --single edges_per_block=128,unroll=2 --iters 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
# L2 probes with torch kernels + unroll variants of the gather probe
python - <<'PY' > gpurun_out/l2probe.log 2>&1
import torch, time
dev=torch.device('cuda')
def t(fn,it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for mb in (8,16,32,48,64,96,128,256,1024):
    x=torch.rand(mb*1024*1024//4,device=dev); y=torch.empty_like(x)
    ms=t(lambda: x.sum()); print('sum',mb,'MB',round(mb/1024/ms*1e3,1),'GB/s... TB/s=',round(mb*1.048576/ms/1e3,2))
    ms=t(lambda: y.copy_(x)); print('copy',mb,'MB r+w TB/s=',round(2*mb*1.048576/ms/1e3,2))
PY
cat gpurun_out/l2probe.log
