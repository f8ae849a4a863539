"""Loader-side plan of the PGCN path: what GPU/PGCN.py:37-64 and :175-182 compute, vectorised.

The reference walks every nnz in interpreted Python on every rank (GPU/PGCN.py:41-45) and keeps an
n x n COO with global indices (:53-64). Here the same information is produced with O(nnz) NumPy:

  * `compute_communication_maps` / `get_partition_of_adjacency_matrix`: same names, arguments and
    return meaning as the reference functions (dicts of sorted global ids per peer; the owned rows
    of A) — drop-in for callers that want the reference's data structures;
  * `build_local_plan`: the compact per-rank layout the B200 kernels consume — local CSR (int32)
    over the column space [own | halo grouped by source peer, sorted by global id], its transpose,
    send_idx / send_off / recv_off. Sender order == receiver order because both sides sort by
    global id (the invariant GPU/PGCN.py:47-48 relies on);
  * `PgcnPlan`: owns the device-side plan (C-ABI handle) and the host stats the reference keeps in
    device counters (GPU/PGCN.py:78-83).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import cabi


def _coo(A):
    A = A.tocoo() if sp.issparse(A) else sp.coo_matrix(A)
    return A


def compute_communication_maps(A, partvec, rank, size):
    """Vectorised GPU/PGCN.py:37-51.

    Returns (send_map, recv_map): dicts keyed by every OTHER rank (possibly empty arrays), values
    sorted int64 global vertex ids. recv_map[p] = columns owned by p referenced by my rows;
    send_map[p] = my columns referenced by rows of p. send_map_r[p] == recv_map_p[r].
    """
    A = _coo(A)
    pv = np.asarray(partvec, dtype=np.int64)
    n = A.shape[0]
    prow, pcol = pv[A.row], pv[A.col]
    cross = prow != pcol
    # I receive column j from part[j] when one of my rows references it
    mine = cross & (prow == rank)
    rkeys = np.unique(pcol[mine] * n + A.col[mine].astype(np.int64))
    # I send my column j to part[i] for every foreign row i that references it
    theirs = cross & (pcol == rank)
    skeys = np.unique(prow[theirs] * n + A.col[theirs].astype(np.int64))
    send_map, recv_map = {}, {}
    for p in range(size):
        if p == rank:
            continue
        lo, hi = np.searchsorted(rkeys, [p * n, (p + 1) * n])
        recv_map[p] = rkeys[lo:hi] - p * n
        lo, hi = np.searchsorted(skeys, [p * n, (p + 1) * n])
        send_map[p] = skeys[lo:hi] - p * n
    return send_map, recv_map


def get_partition_of_adjacency_matrix(A, partvec, rank):
    """Vectorised GPU/PGCN.py:53-64: the entries of A whose ROW is owned by `rank`, global indices,
    global shape, duplicates kept. Returns scipy COO (float32)."""
    A = _coo(A)
    pv = np.asarray(partvec, dtype=np.int64)
    keep = pv[A.row] == rank
    return sp.coo_matrix((A.data[keep].astype(np.float32), (A.row[keep], A.col[keep])), shape=A.shape)


# the reference's spelling (GPU/PGCN.py:53)
get_partitiont_of_adjacency_matrix = get_partition_of_adjacency_matrix


class LocalPlan:
    """Host-side arrays of one rank (all NumPy, no device state)."""

    def __init__(self):
        self.n = 0; self.k = 1; self.rank = 0
        self.m = 0; self.h = 0; self.S = 0
        self.owned = None        # int64[m] global ids of my rows, ascending
        self.halo = None         # int64[h] global ids of halo rows, [peer0 | peer1 | ...], ascending inside
        self.rowptr = self.colidx = self.vals = None
        self.t_rowptr = self.t_colidx = self.t_vals = None
        self.send_idx = None     # int32[S] local row ids
        self.send_gid = None     # int64[S] global ids (wire order)
        self.send_off = None     # int64[k+1]
        self.recv_off = None     # int64[k+1]

    # --- the reference's views -----------------------------------------------------------------
    def send_map(self):
        return {p: self.send_gid[self.send_off[p]:self.send_off[p + 1]] for p in range(self.k) if p != self.rank}

    def recv_map(self):
        return {p: self.halo[self.recv_off[p]:self.recv_off[p + 1]] for p in range(self.k) if p != self.rank}

    def nnz(self):
        return int(self.rowptr[-1])


def build_local_plan(A, partvec, rank, size):
    """Compact per-rank layout (see module docstring). O(nnz) NumPy + scipy COO->CSR."""
    A = _coo(A)
    n = A.shape[0]
    pv = np.asarray(partvec, dtype=np.int64)
    if pv.shape[0] != n:
        raise ValueError("part vector has %d entries, matrix has %d rows" % (pv.shape[0], n))
    if pv.size and (pv.min() < 0 or pv.max() >= size):
        raise KeyError(int(pv.max() if pv.max() >= size else pv.min()))   # the reference's failure mode

    lp = LocalPlan()
    lp.n, lp.k, lp.rank = n, size, rank
    owned = np.flatnonzero(pv == rank).astype(np.int64)
    lp.owned = owned
    lp.m = m = owned.shape[0]
    g2l = np.full(n, -1, dtype=np.int64)
    g2l[owned] = np.arange(m)

    prow, pcol = pv[A.row], pv[A.col]
    mine = prow == rank
    grow, gcol = A.row[mine].astype(np.int64), A.col[mine].astype(np.int64)
    val = A.data[mine].astype(np.float32)
    pc = pcol[mine]

    # halo = distinct foreign columns, grouped by owner then sorted by global id
    foreign = pc != rank
    hkeys = np.unique(pc[foreign] * n + gcol[foreign])
    hpart, hgid = hkeys // n, hkeys % n
    lp.halo = hgid
    lp.h = h = hgid.shape[0]
    lp.recv_off = np.concatenate([[0], np.cumsum(np.bincount(hpart, minlength=size))]).astype(np.int64)

    lcol = g2l[gcol]
    if h:
        lcol[foreign] = m + np.searchsorted(hkeys, pc[foreign] * n + gcol[foreign])
    lrow = g2l[grow]

    csr = sp.coo_matrix((val, (lrow, lcol)), shape=(m, m + h)).tocsr()   # duplicates summed (fp32)
    csr.sort_indices()
    lp.rowptr = csr.indptr.astype(np.int32)
    lp.colidx = csr.indices.astype(np.int32)
    lp.vals = csr.data.astype(np.float32)
    csc = csr.T.tocsr()
    csc.sort_indices()
    lp.t_rowptr = csc.indptr.astype(np.int32)
    lp.t_colidx = csc.indices.astype(np.int32)
    lp.t_vals = csc.data.astype(np.float32)

    # send lists: my columns referenced by rows of peer p, grouped by p, sorted by global id
    theirs = (pcol == rank) & (prow != rank)
    skeys = np.unique(prow[theirs] * n + A.col[theirs].astype(np.int64))
    spart, sgid = skeys // n, skeys % n
    lp.send_gid = sgid
    lp.send_idx = g2l[sgid].astype(np.int32)
    lp.S = int(sgid.shape[0])
    lp.send_off = np.concatenate([[0], np.cumsum(np.bincount(spart, minlength=size))]).astype(np.int64)
    return lp


PLAN_FORMAT_VERSION = 2      # bump when LocalPlan's layout or build_local_plan's ordering / duplicate semantics change

_LP_ARRAYS = ("owned", "halo", "rowptr", "colidx", "vals", "t_rowptr", "t_colidx", "t_vals",
              "send_idx", "send_gid", "send_off", "recv_off")


def save_local_plan(path, lp):
    """Binary cache of one rank's plan: the reference re-reads the .mtx and re-walks
    every nnz on every rank at every start (GPU/PGCN.py:171-176)."""
    np.savez(path, meta=np.array([lp.n, lp.k, lp.rank, lp.m, lp.h, lp.S], dtype=np.int64),
             **{name: getattr(lp, name) for name in _LP_ARRAYS})


def load_local_plan(path):
    z = np.load(path)
    lp = LocalPlan()
    lp.n, lp.k, lp.rank, lp.m, lp.h, lp.S = [int(x) for x in z["meta"]]
    for name in _LP_ARRAYS:
        setattr(lp, name, z[name])
    return lp


def cached_local_plan(path_A, path_partvec, rank, size, cache_dir):
    """build_local_plan with an on-disk cache keyed by the two input files (size + mtime), rank and size."""
    import hashlib
    import os
    from . import graphio
    sa, sp_ = os.stat(path_A), os.stat(path_partvec)
    key = hashlib.sha1(repr([PLAN_FORMAT_VERSION, os.path.abspath(path_A), sa.st_size, sa.st_mtime_ns,
                             os.path.abspath(path_partvec), sp_.st_size, sp_.st_mtime_ns,
                             rank, size]).encode()).hexdigest()[:16]
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, "plan_%s_r%dof%d.npz" % (key, rank, size))
    if os.path.exists(path):
        lp = load_local_plan(path)
        if lp.k == size and lp.rank == rank:
            return lp
    A = graphio.read_adjacency(path_A)
    pv = graphio.check_partvec(graphio.read_partvec(path_partvec, A.shape[0]), size)
    lp = build_local_plan(A, pv, rank, size)
    tmp = path + ".%d.tmp.npz" % os.getpid()
    save_local_plan(tmp, lp)
    os.replace(tmp, path)
    return lp


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class PgcnPlan:
    """Device-side plan = the `A` handle handed to PSpMM.apply.

    Wraps pgcn_plan_create/destroy and carries the host counters the reference keeps as device
    tensors (GPU/PGCN.py:78-83, :105-106, :113-114): volumes in ROWS, message counts including the
    zero-length ones (the reference sends them, GPU/PGCN.py:101-107).
    """

    def __init__(self, local_plan, f_max, device=None):
        import torch
        self.lp = lp = local_plan
        self.f_max = int(f_max)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = cabi.load()
        self._h = C.c_void_p()
        self.layout = "local"
        self.stats = {"send_volume": 0, "recv_volume": 0, "send_nmsg": 0, "recv_nmsg": 0}
        for name in ("rowptr", "colidx", "vals", "t_rowptr", "t_colidx", "t_vals", "send_idx", "send_off", "recv_off"):
            setattr(lp, name, np.ascontiguousarray(getattr(lp, name)))
        with torch.cuda.device(self.device):
            rc = self._lib.pgcn_plan_create(
                _ptr(lp.rowptr), _ptr(lp.colidx), _ptr(lp.vals), lp.m, lp.h,
                _ptr(lp.t_rowptr), _ptr(lp.t_colidx), _ptr(lp.t_vals),
                _ptr(lp.send_idx), _ptr(lp.send_off), _ptr(lp.recv_off),
                lp.k, lp.rank, self.f_max, C.byref(self._h))
        cabi.check(rc, None)
        self._owned_t = None

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pgcn_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h.value:
            raise RuntimeError("plan is closed")
        return self._h

    @property
    def m(self):
        return self.lp.m

    @property
    def n(self):
        return self.lp.n

    def owned_index(self):
        import torch
        if self._owned_t is None:
            self._owned_t = torch.from_numpy(self.lp.owned).to(self.device)
        return self._owned_t

    # -- options / info ------------------------------------------------------------------------
    def set_option(self, name, value):
        cabi.check(self._lib.pgcn_plan_set_option(self.handle, name.encode(), int(value)), self._h)

    def get_option(self, name):
        return int(self._lib.pgcn_plan_get_option(self.handle, name.encode()))

    def autotune(self, f):
        """Pick the fastest edges_per_block for this matrix at width f (set-up work). Returns it."""
        import torch
        with torch.cuda.device(self.device):
            return cabi.check(self._lib.pgcn_plan_autotune(self.handle, int(f)), self._h)

    def algorithmic_bytes(self, f):
        b = cabi.PgcnBytes()
        cabi.check(self._lib.pgcn_algorithmic_bytes(self.handle, int(f), C.byref(b)), self._h)
        return b.as_dict()

    def launch_count(self):
        return int(self._lib.pgcn_launch_count(self.handle))

    # -- communicator --------------------------------------------------------------------------
    def init_comm(self, group=None, transport="auto", nccl_fallback=True):
        """Collective. transport: "nccl" (grouped ncclSend/ncclRecv), "p2p" (peer-memory stores over
        NVLink, single box), or "auto" (p2p when every rank can export/import, else nccl)."""
        import torch
        import torch.distributed as dist
        if self.lp.k == 1:
            return "none"
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before PgcnPlan.init_comm")
        # collectives of the set-up phase run on whatever backend the process group has
        cdev = self.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
        used = None
        if transport in ("p2p", "auto"):
            blob = C.create_string_buffer(cabi.P2P_HANDLE_BYTES)
            with torch.cuda.device(self.device):
                rc = self._lib.pgcn_p2p_export(self.handle, blob)
            ok = torch.tensor([1 if rc == 0 else 0], device=cdev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                mine = torch.frombuffer(bytearray(blob.raw), dtype=torch.uint8).to(cdev)
                allb = [torch.empty_like(mine) for _ in range(self.lp.k)]
                dist.all_gather(allb, mine, group=group)
                packed = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allb)
                with torch.cuda.device(self.device):
                    rc = self._lib.pgcn_p2p_import(self.handle, packed)
                ok = torch.tensor([1 if rc == 0 else 0], device=cdev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 1:
                    dist.barrier(group=group)
                    used = "p2p"
            if used is None:
                # not unanimous: a rank whose own import succeeded must not keep storing into peer slabs while
                # the others talk NCCL (the job would hang) — switch the peer transport off everywhere
                self.set_option("p2p", 0)
                if transport == "p2p":
                    cabi.check(rc if rc < 0 else -5, self._h)
        # the NCCL communicator is always created: it is the transport of the step-by-step entry points
        # (pgcn_exchange) and the fallback of the fused ones for widths the peer-store kernels do not take
        # (f % 4 != 0)
        if used == "p2p" and not nccl_fallback:
            # peer-memory only (e.g. several ranks sharing one device, where NCCL cannot be set up): widths the
            # peer-store kernels do not take (f % 4 != 0) then have no transport and raise
            return used
        ident = torch.zeros(cabi.NCCL_ID_BYTES, dtype=torch.uint8)
        if self.lp.rank == 0:
            buf = C.create_string_buffer(cabi.NCCL_ID_BYTES)
            cabi.check(self._lib.pgcn_comm_unique_id(buf), None)
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        ident = ident.to(cdev)
        dist.broadcast(ident, src=0, group=group)
        raw = bytes(ident.cpu().numpy().tobytes())
        with torch.cuda.device(self.device):
            cabi.check(self._lib.pgcn_comm_init(self.handle, raw), self._h)
        return used or "nccl"

    def share_comm(self, owner):
        """Borrow `owner`'s NCCL communicator (one plan per mini-batch over one process group)."""
        cabi.check(self._lib.pgcn_comm_share(self.handle, owner.handle), self._h)
        self._comm_owner = owner          # keep it alive

    # -- stats, as the reference counts them (rows, messages incl. empty ones) -------------------
    def count_exchange(self, backward=False):
        lp = self.lp
        out_rows = lp.h if backward else lp.S
        in_rows = lp.S if backward else lp.h
        self.stats["send_volume"] += int(out_rows)
        self.stats["recv_volume"] += int(in_rows)
        self.stats["send_nmsg"] += lp.k - 1
        self.stats["recv_nmsg"] += lp.k - 1


def link_local_plans(plans):
    """All ranks' plans live in THIS process (one GPU or several): wire their peer-memory transports to each other
    directly — export every arena, import the k blobs into every plan (same-process peers are reached through plain
    device pointers, no IPC). Used by the single-box tests and by smoke(); a multi-process job uses init_comm."""
    import torch
    k = len(plans)
    blobs = []
    for p in plans:
        blob = C.create_string_buffer(cabi.P2P_HANDLE_BYTES)
        with torch.cuda.device(p.device):
            cabi.check(p._lib.pgcn_p2p_export(p.handle, blob), p._h)
        blobs.append(blob.raw)
    packed = b"".join(blobs)
    for p in plans:
        assert p.lp.k == k
        with torch.cuda.device(p.device):
            cabi.check(p._lib.pgcn_p2p_import(p.handle, packed), p._h)
    return "p2p"


def build_plan(A, partvec, rank, size, f_max, device=None):
    """a1 + a2 + buffer allocation of the reference's `run` (GPU/PGCN.py:175-182) in one call."""
    return PgcnPlan(build_local_plan(A, partvec, rank, size), f_max, device=device)
