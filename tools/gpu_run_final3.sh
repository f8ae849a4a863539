#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/pytest_gpu_final3.log 2>&1
tail -2 gpurun_out/pytest_gpu_final3.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
cut -c1-1200 gpurun_out/bench_n1_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_final.log 2>&1
grep -c spmm_rowblock gpurun_out/launches_final.csv
