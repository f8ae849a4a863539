#!/usr/bin/env python
"""Produce the benchmark part vectors (offline, CPU): bench_data/<config>.<k>.<hp|gp>.npz.

    python tools/make_partvecs.py --config C5 --method hp --k 2 4 8 --preset speed
    python tools/make_partvecs.py --config C2 --method gp --k 8

The partitioning itself is done by oracle/_ref/part_tool_{hp,gp} (oracle/part_tool.cpp): the reference's own
PaToH / METIS libraries, with the models and parameters of the reference drivers GPU/hypergraph/main.cpp:312-386
and GPU/graph/main.cpp:300-360, on the A+I-normalised pattern (partition the matrix WITH its
diagonal, as the reference pipeline does). bench.py --partition auto picks the files up on the GPU box; the .npz
holds the vector as uint8 plus the tool's own report (cut, halo rows per part).
"""
import argparse, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--method", default="hp", choices=["hp", "gp"])
    ap.add_argument("--k", type=int, nargs="+", default=[2, 4, 8])
    ap.add_argument("--preset", default="quality", choices=["quality", "speed", "default"])
    ap.add_argument("--cache", default="/tmp/pgcn_b200_cache")
    ap.add_argument("--out", default=os.path.join(ROOT, "bench_data"))
    ap.add_argument("--parallel", action="store_true", help="run the k values concurrently (one core each)")
    args = ap.parse_args()
    from pgcn_b200 import graphio
    tool = os.path.join(ROOT, "oracle", "_ref", "part_tool_%s" % args.method)
    if not os.path.exists(tool):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "part_tools"])
    os.makedirs(args.out, exist_ok=True)
    csr_path = os.path.join(args.cache, "%s_csr.bin" % args.config)
    A = None
    if not os.path.exists(csr_path):
        A = graphio.config_graph(args.config, cache_dir=args.cache).tocsr()
        A.sort_indices()
        with open(csr_path + ".tmp", "wb") as f:
            np.array([A.shape[0], A.nnz], dtype=np.int64).tofile(f)
            A.indptr.astype(np.int32).tofile(f)
            A.indices.astype(np.int32).tofile(f)
        os.replace(csr_path + ".tmp", csr_path)
    n = graphio.CONFIGS[args.config][0]
    procs = []
    for k in args.k:
        raw = os.path.join(args.cache, "%s.%d.%s.u8" % (args.config, k, args.method))
        log = raw + ".log"
        cmd = [tool, args.method, str(k), csr_path, raw, args.preset, "1"]
        t0 = time.time()
        pr = subprocess.Popen(cmd, stdout=open(log, "w"), stderr=subprocess.STDOUT)
        procs.append((k, raw, log, pr, t0))
        if not args.parallel:
            pr.wait()
    for k, raw, log, pr, t0 in procs:
        rc = pr.wait()
        dt = time.time() - t0
        report = open(log).read()
        if rc != 0:
            print("k=%d FAILED rc=%d\n%s" % (k, rc, report)); continue
        pv = np.fromfile(raw, dtype=np.uint8)
        assert pv.shape[0] == n and pv.max() < k
        out = os.path.join(args.out, "%s.%d.%s.npz" % (args.config, k, args.method))
        np.savez_compressed(out, partvec=pv, report=np.array(report), preset=np.array(args.preset),
                            seconds=np.array(dt), tool=np.array("oracle/part_tool.cpp"))
        print("k=%d %s %.0fs -> %s (%d bytes)\n%s" % (k, args.method, dt, out, os.path.getsize(out), report), flush=True)


if __name__ == "__main__":
    main()
