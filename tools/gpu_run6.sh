#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu6.log 2>&1
tail -4 gpurun_out/pytest_gpu6.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1_v3.json 2> gpurun_out/bench_n1_v3.err
cut -c1-1800 gpurun_out/bench_n1_v3.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-600 gpurun_out/bench_ref.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rowblock -s 3 -c 2 -f -o gpurun_out/prof_c2_v3 \
    python tools/tune_spmm.py --config C2 --single edges_per_block=128,unroll=2 --iters 3 > gpurun_out/ncu_v3.log 2>&1
tail -2 gpurun_out/ncu_v3.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v3.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_v3.log 2>&1
grep -c spmm_rowblock gpurun_out/launches_v3.csv
timeout 300 python tools/tune_spmm.py --config C2 --transpose --sweep mini 2>/dev/null | cut -c1-200
