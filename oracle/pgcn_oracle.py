"""CPU oracle of the PGCN aggregation path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product path (pgcn_b200) never does and has no CPU fallback.

What it restates (plain NumPy / SciPy CSR in fp32, fp64 ground truth beside it; small pure-Python
loops only where the reference itself loops), each function citing the reference lines it follows:

    compute_communication_maps     GPU/PGCN.py:37-51
    partition_of_adjacency_matrix  GPU/PGCN.py:53-64
    communicate_fgm                GPU/PGCN.py:85-119   (all k ranks simulated in one process)
    pspmm_forward / pspmm_backward GPU/PGCN.py:121-134
    grb_aggregate (C, spmm_oracle.c)   Parallel-GCN/main.c:269-299 (fwd) / :374-404 (bwd)

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py imports the unmodified
/root/reference/GPU/PGCN.py, runs its compute_communication_maps / PSpMM under gloo on k = 1, 2, 3
ranks on the shipped fixtures (gemat11 + .3.hp, karate) and commits the results as
tests/golden/*.npz; tests/test_oracle_golden.py checks every function here against them.
The GraphBLAS twin (Parallel-GCN) cannot be built here (no GraphBLAS.h / mpicc): for that boundary
parity is unpinned by the reference and rests on 1-rank equivalence with the PGCN.py oracle.

Quirk model: `literal=True` reproduces the reference bit-for-behaviour —
Q1 (H + X doubles halo rows when H is populated there), Q2 (X is never cleared), Q3 (backward
ASSIGNS received rows: last writer wins, in the reference's receive order). `literal=False` is the
intended semantics the product implements (halo rows read once, backward contributions summed).
"""
import numpy as np
import scipy.sparse as sp


# ---------------------------------------------------------------------------------------------
# a1 / a2: plan inputs
# ---------------------------------------------------------------------------------------------

def compute_communication_maps(A, partvec, rank, size):
    """GPU/PGCN.py:37-51, the same per-nnz loop (use on small inputs). Returns dicts of sorted lists,
    own key popped, every other rank present."""
    A = A.tocoo()
    send_map = {p: set() for p in range(size)}
    recv_map = {p: set() for p in range(size)}
    for i in range(A.nnz):
        r, c = int(A.row[i]), int(A.col[i])
        if partvec[r] == rank and partvec[c] != rank:
            recv_map[partvec[c]].add(c)
        if partvec[c] == rank and partvec[r] != rank:
            send_map[partvec[r]].add(c)
    send_map = {p: np.array(sorted(send_map[p]), dtype=np.int64) for p in range(size)}
    recv_map = {p: np.array(sorted(recv_map[p]), dtype=np.int64) for p in range(size)}
    recv_map.pop(rank)
    send_map.pop(rank)
    return send_map, recv_map


def partition_of_adjacency_matrix(A, partvec, rank):
    """GPU/PGCN.py:53-64: rows owned by `rank`, global shape, uncoalesced (duplicates kept)."""
    A = A.tocoo()
    pv = np.asarray(partvec)
    indices = np.flatnonzero(pv == rank)
    keep = np.isin(A.row, indices)
    return sp.coo_matrix((A.data[keep].astype(np.float32), (A.row[keep], A.col[keep])), shape=A.shape)


# ---------------------------------------------------------------------------------------------
# a3: the exchange, all ranks simulated in-process
# ---------------------------------------------------------------------------------------------

def recv_order(rank, peers):
    """Order in which the reference posts its blocking recvs (GPU/PGCN.py:99-115): phase 0 receives
    from sources with `not rank < source` i.e. lower ranks, iterated in descending order; phase 1
    from higher ranks, descending."""
    lower = sorted([p for p in peers if p < rank], reverse=True)
    higher = sorted([p for p in peers if p > rank], reverse=True)
    return lower + higher


class RankState:
    """The module globals of one reference process (GPU/PGCN.py:23-35, a7): maps, X scratch, stats."""

    def __init__(self, A, partvec, rank, size, f):
        self.rank, self.size = rank, size
        self.send_map, self.recv_map = compute_communication_maps_fast(A, partvec, rank, size)
        self.A = partition_of_adjacency_matrix(A, partvec, rank).tocsr()
        self.X = np.zeros((A.shape[0], f), dtype=np.float32)          # GPU/PGCN.py:189, never re-zeroed
        self.stats = {"send_volume": 0, "recv_volume": 0, "send_nmsg": 0, "recv_nmsg": 0}


def compute_communication_maps_fast(A, partvec, rank, size):
    """Vectorised twin of compute_communication_maps for larger oracle inputs (checked equal to the
    loop version in tests/test_oracle_golden.py)."""
    A = A.tocoo()
    pv = np.asarray(partvec, dtype=np.int64)
    n = A.shape[0]
    pr, pc = pv[A.row], pv[A.col]
    send_map, recv_map = {}, {}
    for p in range(size):
        if p == rank:
            continue
        recv_map[p] = np.unique(A.col[(pr == rank) & (pc == p)]).astype(np.int64)
        send_map[p] = np.unique(A.col[(pc == rank) & (pr == p)]).astype(np.int64)
    return send_map, recv_map


def communicate_fgm(states, Hs, backward=False, literal=True):
    """GPU/PGCN.py:85-119 for all ranks at once. `Hs[r]` is rank r's n x f tensor.

    literal=True : X_r[recv[p]] = H_p[send_p[r]] (ASSIGN, reference receive order), return H + X
                   with the persistent X of each RankState.
    literal=False: intended semantics — fresh zero scratch, received rows ADDED (matters only in
                   backward where a row can arrive from several peers).
    """
    k = len(states)
    out = []
    for r in range(k):
        st = states[r]
        send = st.recv_map if backward else st.send_map
        recv = st.send_map if backward else st.recv_map
        X = st.X if literal else np.zeros_like(Hs[r])
        for src in recv_order(r, list(recv.keys())):
            # what `src` sends me: rows its own send-side map lists for me, in ITS sorted order
            src_send = states[src].recv_map if backward else states[src].send_map
            buf = Hs[src][src_send[r]]
            idx = recv[src]
            assert idx.shape[0] == buf.shape[0]
            if literal:
                X[idx] = buf                                  # GPU/PGCN.py:115
            else:
                np.add.at(X, idx, buf)
            st.stats["recv_volume"] += len(idx)
            st.stats["recv_nmsg"] += 1
        for tgt in send.keys():
            st.stats["send_volume"] += len(send[tgt])          # GPU/PGCN.py:105-106
            st.stats["send_nmsg"] += 1
        out.append(Hs[r] + X)                                  # GPU/PGCN.py:117
    return out


# ---------------------------------------------------------------------------------------------
# a4 / a5: the operator
# ---------------------------------------------------------------------------------------------

def pspmm_forward(states, Hs, literal=True):
    """GPU/PGCN.py:123-127 on every rank: A_r @ communicate_fgm(H_r). fp32 CSR x dense."""
    Hx = communicate_fgm(states, Hs, backward=False, literal=literal)
    return [np.asarray(states[r].A @ Hx[r], dtype=np.float32) for r in range(len(states))]


def pspmm_backward(states, Gs, literal=True):
    """GPU/PGCN.py:129-134 on every rank: communicate_fgm(A_r^T @ g_r, backward=True)."""
    T = [np.asarray(states[r].A.T.tocsr() @ Gs[r], dtype=np.float32) for r in range(len(states))]
    return communicate_fgm(states, T, backward=True, literal=literal)


def truth_forward(A, H):
    """fp64 ground truth of one aggregation on the whole graph: A @ H."""
    return np.asarray(A.tocsr().astype(np.float64) @ H.astype(np.float64))


def truth_backward(A, G):
    """fp64 ground truth of the gradient: A^T @ G."""
    return np.asarray(A.tocsr().astype(np.float64).T @ G.astype(np.float64))


def abs_bound(A, H):
    """|A| @ |H| in fp64 — the scale of the fp32 reassociation error bound of SURVEY.md §8a:
    |Z - Z64| <= 2 * gamma_d * (|A| |H|),  gamma_d ~ d * 2^-24."""
    A = A.tocsr().astype(np.float64)
    return np.asarray(abs(A) @ np.abs(H.astype(np.float64)))


def row_degree(A):
    return np.diff(A.tocsr().indptr)
