"""Mini-batch loader (host side) against the unmodified reference functions
(tests/golden/karate_minibatch.npz, make_minibatch_golden.py): same sampling sequence, same induced
sub-matrices, same per-batch communication maps."""
import os

import numpy as np
import scipy.sparse as sp

from helpers import GOLDEN
from pgcn_b200 import minibatch


def test_batches_equal_reference():
    z = np.load(os.path.join(GOLDEN, "karate_minibatch.npz"))
    n, bs, nb = int(z["n"]), int(z["batch_size"]), int(z["nbatches"])
    A = sp.coo_matrix((z["val"], (z["row"], z["col"])), shape=(n, n))
    pv = z["partvec"].astype(np.int64)
    sets = minibatch.batch_index_sets(n, bs, seed=1)
    assert len(sets) == nb == (n // bs + 1) * 3
    for b in range(nb):
        assert np.array_equal(sets[b], z["b%d_idx" % b])                         # random.sample sequence, seed 1
        bA = minibatch.sample_adjacency_matrix(A, sets[b])
        assert bA.shape == (n, n)
        assert np.array_equal(bA.row, z["b%d_row" % b]) and np.array_equal(bA.col, z["b%d_col" % b])
    for rank in range(3):
        plans, _ = minibatch.batch_local_plans(A, pv, rank, 3, bs, seed=1)
        for b, lp in enumerate(plans):
            assert lp.m == int((pv == rank).sum())                              # every batch keeps the rank's row layout
            for p in range(3):
                if p == rank:
                    continue
                assert np.array_equal(lp.send_map()[p], z["b%d_r%d_send_%d" % (b, rank, p)])
                assert np.array_equal(lp.recv_map()[p], z["b%d_r%d_recv_%d" % (b, rank, p)])
            # the batch plan multiplies like the induced sub-matrix
            H = np.arange(n * 2, dtype=np.float64).reshape(n, 2)
            Aloc = sp.csr_matrix((lp.vals.astype(np.float64), lp.colidx, lp.rowptr), shape=(lp.m, lp.m + lp.h))
            want = (sp.csr_matrix(minibatch.sample_adjacency_matrix(A, sets[b])) @ H)[lp.owned]
            np.testing.assert_allclose(Aloc @ H[np.concatenate([lp.owned, lp.halo])], want)


def test_partvec_pickle_reader(tmp_path):
    import pickle
    path = str(tmp_path / "pv.pkl")
    pickle.dump([0, 2, 1, 1], open(path, "wb"))
    assert minibatch.read_partvec_pickle(path).tolist() == [0, 2, 1, 1]
