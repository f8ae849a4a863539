"""Host-side plan builder (vectorised GPU/PGCN.py:37-64) against the reference's golden maps and
against the fp64 truth of the product it encodes. CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_cases
from helpers import Golden, assert_close_fp32, fp32_tol
from oracle import pgcn_oracle as orc
from pgcn_b200 import graphio, plan as planmod


@pytest.mark.parametrize("case", golden_cases())
def test_maps_equal_reference(case):
    g = Golden(case)
    for r in range(g.k):
        send_ref, recv_ref = g.maps(r)
        send, recv = planmod.compute_communication_maps(g.A, g.partvec, r, g.k)
        lp = planmod.build_local_plan(g.A, g.partvec, r, g.k)
        assert sorted(send) == sorted(send_ref)
        for p in send_ref:
            assert np.array_equal(send[p], send_ref[p]) and np.array_equal(recv[p], recv_ref[p])
            assert np.array_equal(lp.send_map()[p], send_ref[p]) and np.array_equal(lp.recv_map()[p], recv_ref[p])
            assert np.array_equal(lp.owned[lp.send_idx[lp.send_off[p]:lp.send_off[p + 1]]], send_ref[p])


@pytest.mark.parametrize("case", golden_cases())
def test_partition_equals_reference(case):
    g = Golden(case)
    for r in range(g.k):
        P = planmod.get_partitiont_of_adjacency_matrix(g.A, g.partvec, r)
        assert P.shape == (g.n, g.n)
        assert np.array_equal(P.row, g.get(r, "loc_row")) and np.array_equal(P.col, g.get(r, "loc_col"))
        assert np.array_equal(P.data, g.get(r, "loc_val"))


@pytest.mark.parametrize("case", golden_cases())
def test_local_csr_encodes_the_same_product(case):
    g = Golden(case)
    Z64 = orc.truth_forward(g.A, g.H)
    G64 = orc.truth_backward(g.A, g.G)
    acc = np.zeros_like(G64)
    for r in range(g.k):
        lp = planmod.build_local_plan(g.A, g.partvec, r, g.k)
        assert lp.rowptr.dtype == np.int32 and lp.colidx.dtype == np.int32 and lp.vals.dtype == np.float32
        assert lp.rowptr.shape[0] == lp.m + 1 and lp.t_rowptr.shape[0] == lp.m + lp.h + 1
        assert lp.recv_off[-1] == lp.h and lp.send_off[-1] == lp.S
        Aloc = sp.csr_matrix((lp.vals.astype(np.float64), lp.colidx, lp.rowptr), shape=(lp.m, lp.m + lp.h))
        At = sp.csr_matrix((lp.t_vals.astype(np.float64), lp.t_colidx, lp.t_rowptr), shape=(lp.m + lp.h, lp.m))
        assert abs(Aloc.T - At).max() == 0
        cols = np.concatenate([lp.owned, lp.halo])
        np.testing.assert_allclose(Aloc @ g.H[cols].astype(np.float64), Z64[lp.owned], rtol=1e-6, atol=1e-6 * np.abs(Z64).max())
        np.add.at(acc, cols, At @ g.G[lp.owned].astype(np.float64))
    np.testing.assert_allclose(acc, G64, rtol=1e-6, atol=1e-6 * np.abs(G64).max())


def test_edge_cases_empty_rank_and_isolated_rows():
    # 6 vertices, rank 2 owns nothing, vertex 5 has no entries at all, duplicate entry (0,1)
    row = np.array([0, 0, 0, 1, 2, 3, 4, 1])
    col = np.array([1, 1, 3, 0, 4, 3, 0, 2])
    val = np.arange(1, 9, dtype=np.float64)
    A = sp.coo_matrix((val, (row, col)), shape=(6, 6))
    pv = np.array([0, 1, 0, 1, 0, 1])
    H = np.arange(12, dtype=np.float32).reshape(6, 2)
    Z64 = orc.truth_forward(A, H)
    for r in range(3):
        lp = planmod.build_local_plan(A, pv, r, 3)
        if r == 2:
            assert lp.m == 0 and lp.h == 0 and lp.S == 0 and lp.nnz() == 0
            continue
        Aloc = sp.csr_matrix((lp.vals, lp.colidx, lp.rowptr), shape=(lp.m, lp.m + lp.h))
        cols = np.concatenate([lp.owned, lp.halo])
        np.testing.assert_allclose(Aloc @ H[cols], Z64[lp.owned])
        assert sorted(lp.send_map()) == [p for p in range(3) if p != r]       # empty peers keep their key
    with pytest.raises(KeyError):                                           # part id >= size: reference fails the same way
        planmod.build_local_plan(A, np.array([0, 1, 0, 1, 0, 3]), 0, 3)
    with pytest.raises(ValueError):
        planmod.build_local_plan(A, pv[:5], 0, 3)


def test_partvec_io_roundtrip(tmp_path):
    pv = graphio.random_partvec(1000, 4, seed=3)
    path = str(tmp_path / "g.mtx.4.rp")
    graphio.write_partvec(path, pv)
    txt = open(path).read()
    assert txt.endswith(" \n") and txt.count("\n") == 1                       # GPU/hypergraph/main.cpp:51-63 format
    assert np.array_equal(graphio.read_partvec(path, 1000), pv)
    with pytest.raises(ValueError):
        graphio.read_partvec(path, 999)


def test_synthetic_graph_recipe():
    n, nnz = 5000, 60000
    A = graphio.synthetic_graph(n, nnz, seed=1)
    assert A.shape == (n, n) and A.nnz == nnz + n                             # +I adds n self loops
    assert abs(A - A.T).max() < 1e-7                                          # symmetric pattern, symmetric scaling
    B = graphio.synthetic_graph(n, nnz, seed=1)
    assert np.array_equal(A.row, B.row) and np.array_equal(A.data, B.data)    # deterministic
    # normalisation == the reference recipe Dr^-1/2 (A+I) Dc^-1/2 on the 0/1 pattern
    P = sp.csr_matrix((np.ones(A.nnz), (A.row, A.col)), shape=(n, n))
    dr = 1 / np.sqrt(np.asarray(P.sum(1)).ravel()); dc = 1 / np.sqrt(np.asarray(P.sum(0)).ravel())
    np.testing.assert_allclose(A.data, dr[A.row] * dc[A.col], rtol=1e-6)


def test_mtx_loader_matches_scipy(tmp_path):
    from scipy.io import mmwrite
    A = graphio.synthetic_graph(300, 2000, seed=2)
    path = str(tmp_path / "a.mtx")
    mmwrite(path, A)
    B = graphio.read_adjacency(path)
    assert B.shape == A.shape and B.nnz == A.nnz
    graphio.save_adjacency_npz(str(tmp_path / "a.npz"), A)
    Cc = graphio.read_adjacency(str(tmp_path / "a.npz"))
    assert abs(sp.csr_matrix(Cc) - sp.csr_matrix(A)).max() < 1e-7


def test_plan_cache_roundtrip(tmp_path):
    from scipy.io import mmwrite
    g = Golden("gemat11_k3_hp")
    a = str(tmp_path / "g.mtx"); mmwrite(a, g.A, precision=17)
    pv = str(tmp_path / "g.mtx.3.hp"); graphio.write_partvec(pv, g.partvec)
    cache = str(tmp_path / "cache")
    lp1 = planmod.cached_local_plan(a, pv, 1, 3, cache)          # builds and stores
    lp2 = planmod.cached_local_plan(a, pv, 1, 3, cache)          # loads
    ref = planmod.build_local_plan(g.A, g.partvec, 1, 3)
    for name in planmod._LP_ARRAYS:
        assert np.array_equal(getattr(lp1, name), getattr(ref, name)), name
        assert np.array_equal(getattr(lp2, name), getattr(ref, name)), name
        assert getattr(lp2, name).dtype == getattr(ref, name).dtype, name
    assert (lp2.n, lp2.k, lp2.rank, lp2.m, lp2.h, lp2.S) == (ref.n, ref.k, ref.rank, ref.m, ref.h, ref.S)
    assert len([f for f in __import__("os").listdir(cache) if f.endswith(".npz")]) == 1
