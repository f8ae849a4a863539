"""Host side of the mini-batch variant (GPU/PGCN-Mini-batch.py) — SURVEY.md §8f rank 3, loader part only.

The reference pre-samples `nbatches = (n // batch_size + 1) * 3` vertex sets with `random.sample` (seed 1,
GPU/PGCN-Mini-batch.py:201-203,220-230), keeps for each the induced sub-matrix (entries whose row AND column are
in the batch, :58-69), and recomputes the per-peer maps and the rank's rows per batch with the same per-nnz
Python loop as the full-batch trainer (:40-56, :71-82). The same operator then runs on the per-batch matrix.

Here each batch becomes an ordinary LocalPlan (plan.build_local_plan on the induced sub-matrix, global shape and
the global part vector kept), so PSpMM / PgcnPlan run unchanged on it; every batch plan shares the rank's [m, f]
row layout (rows outside the batch are simply empty). The training driver itself is not part of this round.
"""
import pickle
import random

import numpy as np
import scipy.sparse as sp

from . import plan as planmod


def read_partvec_pickle(path):
    """The mini-batch trainer reads a PICKLED list of part ids (GPU/PGCN-Mini-batch.py:216-217), written by
    GPU/SHP/main.py."""
    with open(path, "rb") as f:
        return np.asarray(pickle.load(f), dtype=np.int64)


def sample_adjacency_matrix(A, indices):
    """GPU/PGCN-Mini-batch.py:58-69, vectorised: entries with row in `indices` and column in `indices`;
    shape and entry order of A preserved."""
    A = A.tocoo()
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[np.asarray(indices, dtype=np.int64)] = True
    keep = mask[A.row] & mask[A.col]
    return sp.coo_matrix((A.data[keep], (A.row[keep], A.col[keep])), shape=A.shape)


def batch_index_sets(n, batch_size, seed=1):
    """The reference's sampling sequence: random.seed(seed); (n // batch_size + 1) * 3 draws of
    random.sample(range(n), batch_size) (GPU/PGCN-Mini-batch.py:201-203,220-225)."""
    rnd = random.Random(seed)
    nbatches = (n // batch_size + 1) * 3
    return [np.array(rnd.sample(range(n), batch_size), dtype=np.int64) for _ in range(nbatches)]


def batch_local_plans(A, partvec, rank, size, batch_size, seed=1, index_sets=None):
    """One LocalPlan per pre-sampled batch for this rank (the reference's `batches` list,
    GPU/PGCN-Mini-batch.py:220-230: [bA, batch_send_map, batch_recv_map])."""
    sets = batch_index_sets(A.shape[0], batch_size, seed) if index_sets is None else index_sets
    return [planmod.build_local_plan(sample_adjacency_matrix(A, idx), partvec, rank, size) for idx in sets], sets
