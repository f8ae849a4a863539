#!/usr/bin/env python
"""Does an L2 persisting set-aside (cudaLimitPersistingL2CacheSize) make the evict_last hot rows stick?
Times the C2 forward SpMM at several set-aside sizes in one process."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pgcn_b200 import cabi, graphio, plan as planmod

dev = torch.device("cuda", 0)
torch.cuda.init(); torch.zeros(1, device=dev)
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
n, nnz, f, _, _ = graphio.CONFIGS["C2"]
A = graphio.config_graph("C2", cache_dir="/tmp/pgcn_b200_cache")
lp = planmod.build_local_plan(A, np.zeros(n, dtype=np.int64), 0, 1)
p = planmod.PgcnPlan(lp, f, device=dev)
p.set_option("edges_per_block", 144)
H = torch.rand((n, f), device=dev) * 2 - 1
Z = torch.empty((n, f), device=dev)
lib = cabi.load(); st = torch.cuda.current_stream().cuda_stream
def run():
    cabi.check(lib.pgcn_spmm(p.handle, 0, H.data_ptr(), None, Z.data_ptr(), None, f, st), p.handle)
def timed(it=10):
    for _ in range(3): run()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
prop = torch.cuda.get_device_properties(0)
print(json.dumps({"L2_MB": prop.L2_cache_size / 2**20, "cudart": bool(rt)}), flush=True)
for mb in (0, 32, 48, 64, 80, 0):
    rc = rt.cudaDeviceSetLimit(6, ctypes.c_size_t(mb << 20)) if rt else -1      # cudaLimitPersistingL2CacheSize = 0x06
    got = ctypes.c_size_t(0)
    if rt: rt.cudaDeviceGetLimit(ctypes.byref(got), 6)
    print(json.dumps({"persist_mb": mb, "rc": rc, "limit_now_mb": got.value / 2**20, "ms": timed()}), flush=True)
p.close()
