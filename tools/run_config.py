#!/usr/bin/env python
"""One-off measurement of a benchmark config on one GPU: generate, plan, autotune, time the forward and
transposed aggregation, and check the column-sum checksum (1^T Z == (A^T 1)^T H) at full size."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu-only", action="store_true")
    args = ap.parse_args()
    from pgcn_b200 import graphio, plan as planmod
    n, nnz, f, _, abcd = graphio.CONFIGS[args.config]
    t0 = time.time()
    A = graphio.synthetic_graph(n, nnz, abcd=abcd, seed=1)
    t_gen = time.time() - t0
    t0 = time.time()
    lp = planmod.build_local_plan(A, np.zeros(n, dtype=np.int64), 0, 1)
    t_plan = time.time() - t0
    rec = {"config": args.config, "n": n, "nnz": int(lp.nnz()), "f": f, "gen_s": t_gen, "plan_s": t_plan,
           "max_degree": int(np.diff(lp.rowptr).max())}
    if args.cpu_only:
        print(json.dumps(rec)); return
    import torch
    from pgcn_b200 import cabi
    dev = torch.device("cuda", 0)
    t0 = time.time()
    p = planmod.PgcnPlan(lp, f, device=dev)
    rec["upload_s"] = time.time() - t0
    rec["epb"] = p.autotune(f)
    gen = torch.Generator(device=dev).manual_seed(1)
    H = torch.rand((n, f), device=dev, generator=gen) * 2 - 1
    Z = torch.empty((n, f), device=dev)
    lib = cabi.load(); st = torch.cuda.current_stream().cuda_stream
    ab = p.algorithmic_bytes(f)
    peak = 6567.4
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for tr in (0, 1):
        def run():
            cabi.check(lib.pgcn_spmm(p.handle, tr, H.data_ptr(), None, Z.data_ptr(), None, f, st), p.handle)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        key = "bwd" if tr else "fwd"
        alg = ab["spmm_bwd"] if tr else ab["spmm_fwd"]
        rec[key] = {"ms": ms, "G_edges_per_s": lp.nnz() / ms / 1e6, "alg_GBs": alg / ms / 1e6, "frac": alg / ms / 1e6 / peak,
                    "gather_GBs": ab["gather_fwd"] / ms / 1e6}
    # checksum at full size (forward result is in Z after the last fwd... recompute forward)
    cabi.check(lib.pgcn_spmm(p.handle, 0, H.data_ptr(), None, Z.data_ptr(), None, f, st), p.handle)
    colsum = torch.from_numpy(np.bincount(lp.colidx, weights=lp.vals.astype(np.float64), minlength=n)).to(dev)
    lhs = Z.double().sum(0); rhs = colsum @ H.double()
    rec["checksum_max_rel_err"] = float(((lhs - rhs).abs() / (rhs.abs() + 1e-3)).max())
    rec["algorithmic_bytes_fwd"] = ab["spmm_fwd"]
    print(json.dumps(rec))
    p.close()


if __name__ == "__main__":
    main()
