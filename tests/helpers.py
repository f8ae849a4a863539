"""Shared helpers of the test-suite (golden loading, seeded inputs, tolerances)."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    """One tests/golden/<case>.npz written by make_golden.py from the unmodified reference."""

    def __init__(self, case):
        z = np.load(os.path.join(GOLDEN, case + ".npz"))
        self.z = z
        self.n, self.k, self.f, self.seed = int(z["n"]), int(z["k"]), int(z["f"]), int(z["seed"])
        self.A = sp.coo_matrix((z["val"], (z["row"], z["col"])), shape=(self.n, self.n))
        self.partvec = z["partvec"].astype(np.int64)
        rng = np.random.RandomState(self.seed)            # same draws as make_golden.case_inputs
        self.H = rng.uniform(-1.0, 1.0, size=(self.n, self.f)).astype(np.float32)
        self.G = rng.uniform(-1.0, 1.0, size=(self.n, self.f)).astype(np.float32)

    def owned(self, r):
        return np.flatnonzero(self.partvec == r)

    def masked_H(self, r):
        Hr = self.H.copy()
        Hr[self.partvec != r] = 0.0
        return Hr

    def maps(self, r):
        send = {p: self.z["r%d_send_%d" % (r, p)] for p in range(self.k) if p != r}
        recv = {p: self.z["r%d_recv_%d" % (r, p)] for p in range(self.k) if p != r}
        return send, recv

    def get(self, r, name):
        return self.z["r%d_%s" % (r, name)]


def fp32_tol(A, H, deg_max):
    """Elementwise bound of SURVEY.md §8a: |Z - Z64| <= 2 * gamma_d * (|A| |H|), gamma_d = d * 2^-24,
    plus one ulp-scale slack for the final rounding of the fp32 result."""
    A = sp.csr_matrix(A).astype(np.float64)
    bound = np.asarray(abs(A) @ np.abs(H.astype(np.float64)))
    return 2.0 * (deg_max + 2) * 2.0 ** -24 * bound + 1e-30


def assert_close_fp32(Z, Z64, tol, what=""):
    err = np.abs(Z.astype(np.float64) - Z64)
    bad = err > tol
    assert not bad.any(), "%s: %d entries beyond the fp32 bound, worst %.3e (tol %.3e)" % (
        what, int(bad.sum()), float(err[bad].max()), float(tol[bad].min()))
