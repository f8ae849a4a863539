#!/bin/bash
# One gpurun call for the per-config table (single GPU): forward/backward SpMM time, G edges/s, roofline fraction
# and the full-size checksum for every BASELINE.json config that fits one B200.
#   gpurun --timeout 2400 -- 'bash tools/gpu_configs_table.sh'
mkdir -p gpurun_out
for c in C2 C3 C5 C4; do
  timeout 1200 python tools/run_config.py --config $c --iters 10 > gpurun_out/config_$c.json 2> gpurun_out/config_$c.err
  cat gpurun_out/config_$c.json
done
