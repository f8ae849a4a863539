"""The CPU oracle against the outputs of the unmodified reference (tests/golden, make_golden.py).

This is what pins the oracle (the reference has no tests or golden vectors of its
own, so the vectors are outputs of the reference itself run under gloo in the build container)."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_cases
from helpers import GOLDEN, Golden, assert_close_fp32, fp32_tol
from oracle import build_oracle, pgcn_oracle as orc


@pytest.mark.parametrize("case", golden_cases())
def test_maps_match_reference(case):
    g = Golden(case)
    for r in range(g.k):
        send_ref, recv_ref = g.maps(r)
        loop = orc.compute_communication_maps(g.A, g.partvec, r, g.k) if g.A.nnz < 40000 else None
        fast = orc.compute_communication_maps_fast(g.A, g.partvec, r, g.k)
        for maps in (loop, fast):
            if maps is None:
                continue
            send, recv = maps
            assert sorted(send) == sorted(send_ref) and sorted(recv) == sorted(recv_ref)   # every other rank is a key
            for p in send_ref:
                assert np.array_equal(send[p], send_ref[p])
                assert np.array_equal(recv[p], recv_ref[p])


@pytest.mark.parametrize("case", golden_cases())
def test_partition_matches_reference(case):
    g = Golden(case)
    for r in range(g.k):
        P = orc.partition_of_adjacency_matrix(g.A, g.partvec, r)
        assert P.shape == (g.n, g.n)                                  # global shape kept (GPU/PGCN.py:63)
        assert np.array_equal(P.row, g.get(r, "loc_row"))
        assert np.array_equal(P.col, g.get(r, "loc_col"))
        assert np.array_equal(P.data, g.get(r, "loc_val"))


@pytest.mark.parametrize("case", golden_cases())
def test_forward_backward_literal_match_reference(case):
    """Z1 (Q0 preconditions), the literal gradient (Q2 + Q3) and the second forward with stale X (Q2)."""
    g = Golden(case)
    states = [orc.RankState(g.A, g.partvec, r, g.k, g.f) for r in range(g.k)]
    Hs = [g.masked_H(r) for r in range(g.k)]
    Z1 = orc.pspmm_forward(states, Hs, literal=True)
    Gs = [g.G.copy() for _ in range(g.k)]
    Hg = orc.pspmm_backward(states, Gs, literal=True)
    Z2 = orc.pspmm_forward(states, Hs, literal=True)
    for r in range(g.k):
        own = g.owned(r)
        # same arithmetic up to summation order inside the CSR kernels: compare with a tight flat tol
        for name, mine in (("Z1_own", Z1[r][own]), ("Hgrad_own", Hg[r][own]), ("Z2_own", Z2[r][own])):
            ref = g.get(r, name)
            scale = max(1.0, float(np.abs(ref).max()))
            np.testing.assert_allclose(mine, ref, rtol=2e-5, atol=2e-6 * scale, err_msg="%s r%d %s" % (case, r, name))
        assert np.all(Z1[r][np.setdiff1d(np.arange(g.n), own)] == 0.0)          # non-owned rows exactly 0
        st = states[r].stats
        assert [st["send_volume"], st["recv_volume"], st["send_nmsg"], st["recv_nmsg"]] == list(g.get(r, "stats"))


@pytest.mark.parametrize("case", golden_cases())
def test_intended_semantics_vs_truth_and_reference(case):
    """Intended semantics == fp64 truth within the fp32 bound; == reference where the reference is
    right: forward everywhere (Q0), backward on rows that sit in fewer than two send maps (Q3)."""
    g = Golden(case)
    states = [orc.RankState(g.A, g.partvec, r, g.k, g.f) for r in range(g.k)]
    Hs = [g.masked_H(r) for r in range(g.k)]
    Z = orc.pspmm_forward(states, Hs, literal=False)
    Z64 = orc.truth_forward(g.A, g.H)
    dmax = int(orc.row_degree(g.A).max())
    tolZ = fp32_tol(g.A, g.H, dmax)
    Gs = []
    for r in range(g.k):                       # upstream gradient lives on owned rows only
        Gr = g.G.copy(); Gr[g.partvec != r] = 0.0; Gs.append(Gr)
    Hg = orc.pspmm_backward(states, Gs, literal=False)
    G64 = orc.truth_backward(g.A, g.G)
    tolG = fp32_tol(g.A.T, g.G, int(orc.row_degree(g.A.T).max()))
    for r in range(g.k):
        own = g.owned(r)
        assert_close_fp32(Z[r][own], Z64[own], tolZ[own], "%s fwd r%d" % (case, r))
        np.testing.assert_allclose(Z[r][own], g.get(r, "Z1_own"), rtol=2e-5, atol=2e-6 * max(1.0, np.abs(Z64).max()))
        assert_close_fp32(Hg[r][own], G64[own], tolG[own], "%s bwd r%d" % (case, r))
        send, _ = g.maps(r)
        cnt = np.zeros(g.n, dtype=np.int64)
        for p in send:
            cnt[send[p]] += 1
        single = cnt[own] < 2
        ref = g.get(r, "Hgrad_own")
        np.testing.assert_allclose(Hg[r][own][single], ref[single], rtol=2e-5, atol=2e-6 * max(1.0, np.abs(G64).max()))
        if g.k >= 3 and (~single).any():
            # Q3 is real: on multiply-sent rows the reference differs from the truth somewhere
            assert np.abs(ref[~single] - G64[own][~single]).max() > 1e-3


@pytest.mark.parametrize("case", ["gemat11_k3_hp", "karate_k3_hp", "gemat11_k1"])
def test_c_oracle_matches_numpy_oracle(case):
    """spmm_oracle.c (restating Parallel-GCN/main.c:271,295) == PGCN.py oracle on the same rank data."""
    from pgcn_b200 import plan as planmod
    g = Golden(case)
    Z64 = orc.truth_forward(g.A, g.H)
    dmax = int(orc.row_degree(g.A).max())
    tol = fp32_tol(g.A, g.H, dmax)
    for r in range(g.k):
        lp = planmod.build_local_plan(g.A, g.partvec, r, g.k)
        Hcat = np.concatenate([g.H[lp.owned], g.H[lp.halo]], axis=0)
        Zc = build_oracle.grb_aggregate(lp.rowptr, lp.colidx, lp.vals, Hcat, lp.recv_off, lp.m)
        Zs = build_oracle.spmm_csr(lp.rowptr, lp.colidx, lp.vals, Hcat, lp.m)
        assert_close_fp32(Zc, Z64[lp.owned], tol[lp.owned], "grb_aggregate r%d" % r)
        assert_close_fp32(Zs, Z64[lp.owned], tol[lp.owned], "spmm_csr r%d" % r)
        np.testing.assert_allclose(Zc, g.get(r, "Z1_own"), rtol=2e-5, atol=2e-6 * max(1.0, np.abs(Z64).max()))


def test_reference_e2e_stats_recorded():
    rec = json.load(open(os.path.join(GOLDEN, "gemat11_k3_hp_e2e.json")))
    assert rec["total_vol"] == 36920 and rec["total_nmsg"] == 120          # SURVEY.md §8c probe
    assert len(rec["losses"]) == 4
    # 5 epochs x 2 layers x 2 directions x sum over ranks of rows sent per exchange
    g = Golden("gemat11_k3_hp")
    fwd_rows = sum(len(v) for r in range(3) for v in g.maps(r)[0].values())
    bwd_rows = sum(len(v) for r in range(3) for v in g.maps(r)[1].values())
    assert 5 * 2 * (fwd_rows + bwd_rows) == rec["total_vol"]
