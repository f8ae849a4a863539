"""Input formats of the PGCN path and the synthetic inputs of the benchmark configs.

Formats kept from the reference (SURVEY.md §8b "Input formats"):
  * adjacency: MatrixMarket coordinate file, whatever scipy.io.mmread accepts   GPU/PGCN.py:171
  * part vector: FIRST line of whitespace-separated ints, length n               GPU/PGCN.py:172-173
    (writer: GPU/hypergraph/main.cpp:51-63 — "%d " per vertex then a newline)
Synthetic inputs: R-MAT graph, symmetrised, self-loops removed, then the
reference preprocessing A^ = Dr^-1/2 (A + I) Dc^-1/2                  preprocess/GrB-GNN-IDG.py:45-68
"""
import os

import numpy as np
import scipy.sparse as sp
from scipy.io import mmread


def read_adjacency(path):
    """scipy COO (global n x n), exactly what GPU/PGCN.py:171 hands to the plan builder.
    A `.npz` written by save_adjacency_npz is accepted too (fast path for 1e8-edge graphs)."""
    if path.endswith(".npz"):
        z = np.load(path)
        n = int(z["n"])
        return sp.coo_matrix((z["data"], (z["row"], z["col"])), shape=(n, n))
    A = mmread(path)
    return A.tocoo()


def save_adjacency_npz(path, A):
    A = A.tocoo()
    np.savez(path, n=A.shape[0], row=A.row.astype(np.int32), col=A.col.astype(np.int32),
             data=A.data.astype(np.float32))


def read_partvec(path, n=None):
    """First line, whitespace-separated ints (GPU/PGCN.py:172-173)."""
    with open(path) as f:
        pv = np.array(f.readline().split(), dtype=np.int64)
    if n is not None and pv.shape[0] != n:
        raise ValueError("part vector has %d entries, matrix has %d rows" % (pv.shape[0], n))
    return pv


def write_partvec(path, partvec):
    """Same bytes as GPU/hypergraph/main.cpp:51-63."""
    with open(path, "w") as f:
        f.write("".join("%d " % int(p) for p in partvec))
        f.write("\n")


def check_partvec(partvec, size):
    """The reference fails with KeyError on a part id >= size; same here, earlier."""
    pv = np.asarray(partvec)
    if pv.size and (pv.min() < 0 or pv.max() >= size):
        bad = int(pv.max() if pv.max() >= size else pv.min())
        raise KeyError(bad)
    return pv


# ------------------------------------------------------------------------------------------
# synthetic graphs
# ------------------------------------------------------------------------------------------

def rmat_edges(n, n_undirected, abcd=(0.57, 0.19, 0.19, 0.05), seed=1, permute=True):
    """`n_undirected` distinct undirected edges {u,v}, u != v, of an R-MAT graph on n vertices.

    scale = ceil(log2 n); endpoints >= n are dropped; draws continue until enough distinct edges
    exist, then exactly n_undirected are kept. Vertex ids are randomly permuted (Graph500 style)
    unless permute=False, so no locality comes for free from the generator's bit structure.
    """
    rng = np.random.default_rng(seed)
    scale = max(1, int(np.ceil(np.log2(max(n, 2)))))
    a, b, c, _ = abcd
    keys = np.empty(0, dtype=np.int64)
    want = int(n_undirected)
    draw = int(want * 1.25) + 1024
    while True:
        u = np.zeros(draw, dtype=np.int64)
        v = np.zeros(draw, dtype=np.int64)
        for _ in range(scale):
            r = rng.random(draw)
            ubit = r >= (a + b)
            vbit = ((r >= a) & (r < a + b)) | (r >= a + b + c)
            u = (u << 1) | ubit
            v = (v << 1) | vbit
        ok = (u < n) & (v < n) & (u != v)
        u, v = u[ok], v[ok]
        lo, hi = np.minimum(u, v), np.maximum(u, v)
        keys = np.unique(np.concatenate([keys, lo * n + hi]))
        if keys.shape[0] >= want:
            break
        draw = int((want - keys.shape[0]) * 1.5) + 1024
    if keys.shape[0] > want:
        keys = keys[np.sort(rng.permutation(keys.shape[0])[:want])]
    lo, hi = keys // n, keys % n
    if permute:
        perm = rng.permutation(n)
        lo, hi = perm[lo], perm[hi]
    return lo.astype(np.int64), hi.astype(np.int64)


def symmetric_pattern(n, lo, hi):
    """COO pattern with both directions of every undirected edge (values 1.0, no diagonal)."""
    row = np.concatenate([lo, hi])
    col = np.concatenate([hi, lo])
    return sp.coo_matrix((np.ones(row.shape[0], dtype=np.float32), (row, col)), shape=(n, n))


def gcn_normalise(A):
    """A^ = Dr^-1/2 (A + I) Dc^-1/2 with the diagonal of A dropped first
    (preprocess/GrB-GNN-IDG.py:45-68). Returns float32 COO."""
    A = sp.coo_matrix(A)
    keep = A.row != A.col
    n = A.shape[0]
    row = np.concatenate([A.row[keep], np.arange(n)])
    col = np.concatenate([A.col[keep], np.arange(n)])
    dat = np.concatenate([A.data[keep].astype(np.float64), np.ones(n)])
    B = sp.coo_matrix((dat, (row, col)), shape=(n, n))
    col_sum = np.asarray(B.sum(axis=0)).reshape(-1)
    row_sum = np.asarray(B.sum(axis=1)).reshape(-1)
    dc = 1.0 / np.sqrt(col_sum)
    dr = 1.0 / np.sqrt(row_sum)
    val = (dr[B.row] * B.data * dc[B.col]).astype(np.float32)
    return sp.coo_matrix((val, (B.row.astype(np.int64), B.col.astype(np.int64))), shape=(n, n))


def synthetic_graph(n, nnz, abcd=(0.57, 0.19, 0.19, 0.05), seed=1, permute=True):
    """The benchmark adjacency: R-MAT with `nnz` stored entries before +I (nnz/2 undirected edges),
    symmetrised, normalised like the reference preprocessing. nnz(A^) = nnz + n."""
    lo, hi = rmat_edges(n, nnz // 2, abcd=abcd, seed=seed, permute=permute)
    return gcn_normalise(symmetric_pattern(n, lo, hi))


def random_partvec(n, k, seed=1):
    """`rp` vectors: uniform random parts (GPU/hypergraph/main.cpp:134 partition_random)."""
    if k == 1:
        return np.zeros(n, dtype=np.int64)
    return np.random.default_rng(seed).integers(0, k, size=n, dtype=np.int64)


def block_partvec(n, k):
    """Contiguous equal blocks of vertex ids (a cheap stand-in when no partitioner output exists)."""
    return (np.arange(n, dtype=np.int64) * k // max(n, 1)).astype(np.int64)


CONFIGS = {
    # name: (n, nnz before +I, f, layers, abcd)      BASELINE.json "configs" / SURVEY.md §8d
    "C1": (2708, 10556, 16, 2, (0.57, 0.19, 0.19, 0.05)),
    "C2": (1_000_000, 16_000_000, 128, 2, (0.57, 0.19, 0.19, 0.05)),
    "C3": (2_400_000, 62_000_000, 128, 3, (0.45, 0.22, 0.22, 0.11)),
    "C4": (233_000, 114_000_000, 256, 2, (0.57, 0.19, 0.19, 0.05)),
    "C5": (10_000_000, 100_000_000, 128, 3, (0.57, 0.19, 0.19, 0.05)),
}


def config_graph(name, cache_dir=None, seed=1):
    """Adjacency of a named config, cached as .npz under cache_dir when given."""
    n, nnz, _, _, abcd = CONFIGS[name]
    path = None
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        path = os.path.join(cache_dir, "%s_seed%d.npz" % (name, seed))
        if os.path.exists(path):
            return read_adjacency(path)
    A = synthetic_graph(n, nnz, abcd=abcd, seed=seed)
    if path:
        tmp = path + ".%d.tmp.npz" % os.getpid()
        save_adjacency_npz(tmp, A)
        os.replace(tmp, path)
    return A
