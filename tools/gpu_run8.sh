#!/bin/bash
mkdir -p gpurun_out
# does the per-edge L2 policy change the hit rate at all?  (ld2 variant, hot set 48 MB)  vs default
for cfg in "ld2:48" "ld2:24" ":64"; do
  v=${cfg%%:*}; hot=${cfg#*:}
  export PGCN_B200_VARIANT=$v PGCN_HOT_MB=$hot
  timeout 600 ncu --clock-control none -k regex:spmm_rowblock -s 3 -c 1 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_op_read_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --csv --log-file gpurun_out/ncu_hint_${v}_${hot}.csv python tools/tune_spmm.py --config C2 --single edges_per_block=128 --iters 3 > /dev/null 2>&1
  echo "== $v hot=$hot"; grep -E "dram__bytes|hit_rate|time_duration|inst_executed|issue_active|warps_active|throughput" gpurun_out/ncu_hint_${v}_${hot}.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
  timeout 300 python tools/tune_spmm.py --config C2 --single edges_per_block=128 --iters 10 2>/dev/null | grep edges_per_block | cut -c1-120
done
