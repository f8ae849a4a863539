"""Golden vectors for the mini-batch loader: the UNMODIFIED reference functions of GPU/PGCN-Mini-batch.py
(sample_adjacency_matrix :58-69, compute_communication_maps :40-56, the sampling sequence of run :201-230) on the
shipped karate graph + pickled 3-way part vector.   python tests/golden/make_minibatch_golden.py"""
import importlib.util
import os
import pickle
import random
import sys
import warnings

import numpy as np
import torch
from scipy.io import mmread

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    warnings.filterwarnings("ignore")
    spec = importlib.util.spec_from_file_location("ref_mb", os.path.join(REF, "GPU", "PGCN-Mini-batch.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.device = torch.device("cpu")
    A = mmread(os.path.join(REF, "GPU/SHP/data/karate/karate.mtx")).tocoo()
    partvec = pickle.load(open(os.path.join(REF, "GPU/SHP/data/partvec.hp.3"), "rb"))
    n, batch_size, size = A.shape[0], 12, 3
    random.seed(1)                                            # GPU/PGCN-Mini-batch.py:201-203
    nbatches = (n // batch_size + 1) * 3
    out = {"n": np.array(n), "batch_size": np.array(batch_size), "nbatches": np.array(nbatches),
           "partvec": np.array(partvec, dtype=np.int32), "row": A.row.astype(np.int32), "col": A.col.astype(np.int32),
           "val": A.data.astype(np.float64)}
    for b in range(nbatches):
        idx = np.array(random.sample(range(n), batch_size))    # :224
        bA = ref.sample_adjacency_matrix(A, idx)                # :225
        out["b%d_idx" % b] = idx.astype(np.int32)
        out["b%d_row" % b] = bA.row.astype(np.int32); out["b%d_col" % b] = bA.col.astype(np.int32)
        for rank in range(size):
            ref.myrank = rank
            send, recv = ref.compute_communication_maps(bA, partvec, rank, size)    # :226
            for p in send:
                out["b%d_r%d_send_%d" % (b, rank, p)] = send[p].numpy()
                out["b%d_r%d_recv_%d" % (b, rank, p)] = recv[p].numpy()
    np.savez_compressed(os.path.join(HERE, "karate_minibatch.npz"), **out)
    print("wrote karate_minibatch.npz with", nbatches, "batches")


if __name__ == "__main__":
    sys.exit(main())
