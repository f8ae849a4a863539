"""Fixtures for the CPU-path format reader: run the reference's own partitioner (GCN-HP/main.cpp, compiled
unmodified into oracle/_ref/gcnhgp_cpu by oracle/Makefile) on the shipped karate graph after the reference
preprocessing, and on a small UNSYMMETRIC graph (random partition, -r) that exposes the transposed
connectivity of GCN-HP/main.cpp:154-170. Outputs are committed under tests/golden/cpu_path/.

    make -C oracle _ref/gcnhgp_cpu && python tests/golden/make_cpu_path_fixtures.py
"""
import os
import shutil
import subprocess
import sys

import numpy as np
import scipy.sparse as sp
from scipy.io import mmread, mmwrite

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pgcn_b200 import graphio  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "gcnhgp_cpu")


def run(name, A, k, f, extra=()):
    out = os.path.join(HERE, "cpu_path", name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    n = A.shape[0]
    mmwrite(os.path.join(out, "input.A.mtx"), A, precision=3)          # preprocess/GrB-GNN-IDG.py:80
    mmwrite(os.path.join(out, "input.H.mtx"), sp.coo_matrix(np.ones((n, f))), precision=1)
    Y = np.ones((n, 2)); Y[:, 0] = 0
    mmwrite(os.path.join(out, "input.Y.mtx"), sp.coo_matrix(Y), precision=1)
    cmd = [BIN, "-a", os.path.join(out, "input.A.mtx"), "-h", os.path.join(out, "input.H.mtx"),
           "-y", os.path.join(out, "input.Y.mtx"), "-o", out, "-k", str(k), "-f", str(f), "-l", "3"] + list(extra)
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    for fn in ("input.H.mtx", "input.Y.mtx"):
        os.remove(os.path.join(out, fn))
    for r in range(k):
        os.remove(os.path.join(out, "Y.%d" % r))
    print(name, sorted(os.listdir(out)))


def main():
    kar = mmread("/root/reference/GPU/SHP/data/karate/karate.mtx")
    run("karate_k3", graphio.gcn_normalise(kar), 3, 4)
    rng = np.random.RandomState(5)
    n = 60
    row = rng.randint(0, n, 400); col = rng.randint(0, n, 400)
    val = rng.randint(1, 99, 400) / 100.0
    U = sp.coo_matrix((val, (row, col)), shape=(n, n)).tocsr().tocoo()   # unsymmetric, duplicates merged
    run("unsym_k4_rp", U, 4, 4, extra=("-r",))


if __name__ == "__main__":
    main()
