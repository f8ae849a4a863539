#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
( timeout 600 python -m pytest tests/test_multigpu.py -x -q -k "more_gpus and 8" ) > gpurun_out/pytest_mg8.log 2>&1
tail -4 gpurun_out/pytest_mg8.log
( timeout 600 python -m pytest tests/test_cli.py tests/test_multigpu.py -x -q -m gpu -k "multi_rank or (more_gpus and 3)" ) > gpurun_out/pytest_mg8b.log 2>&1
tail -4 gpurun_out/pytest_mg8b.log
run_bench () {  # n transport partition
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29655 \
     bench.py --gpus $1 --steps 30 --warmup 5 --transport $2 --partition $3 > gpurun_out/bench_n$1_$2_$3.json 2> gpurun_out/bench_n$1_$2_$3.err
  grep '^{' gpurun_out/bench_n$1_$2_$3.json | python -c "
import sys, json
r=json.loads(sys.stdin.read())
print('N=%d %s %s: step %.3f ms -> %.2f G edges/s | kernel %.3f ms | bwd %.3f ms | e2e %.2f ms | halo0 %d rows | xchg %.1f MB' % (r['n_gpus'], r['config']['transport'], r['config']['partition'][:5], r['ms_per_step'], r['value']/1e9, r['roofline']['ms_per_launch'], r['backward']['ms_per_step'], r['e2e']['ms_per_step'], r['config']['halo_rows_rank0'], r['exchange_bytes_in_per_step']/1e6))"
  tail -2 gpurun_out/bench_n$1_$2_$3.err | cut -c1-300
}
run_bench 8 auto block
run_bench 8 nccl block
run_bench 4 auto block
run_bench 2 auto block
