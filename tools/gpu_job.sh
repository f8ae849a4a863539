#!/bin/bash
# The one parameterised GPU-box runner:  gpurun --timeout N -- 'bash tools/gpu_job.sh <job> [args]'
# Everything a job writes goes to gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
job=${1:-tests}; shift || true
case "$job" in
  tests)        # the GPU parity suite
    timeout 1500 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log ;;
  ring-sweep)   # kernel variants on a config (default C2)
    cfg=${1:-C2}; sw=${2:-ring}
    timeout 1500 python tools/tune_spmm.py --config $cfg --sweep $sw --iters 10 --out gpurun_out/ring_sweep_$cfg.jsonl 2> gpurun_out/ring_sweep_$cfg.err | tail -60 ;;
  ncu-spmm)     # one full ncu capture of the SpMM kernel with the given options, e.g. kernel=5,ring_slots=32
    opts=${1:-kernel=5}; tag=${2:-ring}
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 4 -c 2 -f -o gpurun_out/prof_$tag \
        python tools/tune_spmm.py --config C2 --single "$opts" --iters 2 > gpurun_out/ncu_$tag.log 2>&1
    tail -5 gpurun_out/ncu_$tag.log ;;
  traffic)      # DRAM bytes per launch of the default SpMM kernel -> gpurun_out/traffic_<cfg>.json (copy to profiles/)
    cfg=${1:-C2}
    timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:spmm_ring -s 4 -c 3 --csv --log-file gpurun_out/traffic_raw_$cfg.csv \
        python tools/tune_spmm.py --config $cfg --single kernel=0 --iters 3 > gpurun_out/traffic_$cfg.log 2>&1
    python tools/update_traffic.py gpurun_out/traffic_raw_$cfg.csv $cfg gpurun_out/traffic_$cfg.json ;;
  configs)      # single-GPU table of the BASELINE configs (forward / transposed SpMM, checksum)
    for c in "$@"; do
      timeout 1200 python tools/run_config.py --config $c --iters 10 > gpurun_out/config_$c.json 2> gpurun_out/config_$c.err
      cat gpurun_out/config_$c.json
    done ;;
  bench)        # bench.py with the given flags
    timeout 1500 python bench.py "$@" 2> gpurun_out/bench.err | tee gpurun_out/bench_last.json ;;
  launches)     # ncu launch list of a short bench run
    timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
    tail -3 gpurun_out/bench_under_ncu.log ;;
  *) echo "unknown job $job"; exit 2 ;;
esac
