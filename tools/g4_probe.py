#!/usr/bin/env python
"""GPU probe: does the tile::gather4 ring kernel (kernel=7) reproduce the register kernel (kernel=4)?
   python tools/g4_probe.py <g4_box_rows>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pgcn_b200 import cabi, graphio, plan as planmod, op

box = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
for f in (128, 256):
    n = 6000
    A = graphio.synthetic_graph(n, 150000, seed=3)
    for k in (1, 2):
        pv = np.zeros(n, dtype=np.int64) if k == 1 else graphio.random_partvec(n, 2, seed=5)
        p = planmod.build_plan(A, pv, 0, k, f, device=dev)
        lp = p.lp
        H = torch.rand((lp.m, f), device=dev) * 2 - 1
        halo = torch.rand((max(lp.h, 1), f), device=dev) * 2 - 1
        p.set_option("kernel", 4)
        Z4 = op.spmm_local(p, H, halo if lp.h else None).clone()
        p.set_option("kernel", 7); p.set_option("g4_box_rows", box); p.set_option("ring_slots", 16)
        Z7 = op.spmm_local(p, H, halo if lp.h else None)
        torch.cuda.synchronize()
        print("box_rows=%d f=%d k=%d max|diff|=%.3e (max|Z|=%.3e)" % (box, f, k, float((Z4 - Z7).abs().max()), float(Z4.abs().max())), flush=True)
        p.close()
