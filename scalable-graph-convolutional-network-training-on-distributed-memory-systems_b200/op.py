"""PSpMM — the operator boundary of the hot path (GPU/PGCN.py:121-134), B200-native.

    PSpMM.apply(A, H)     A = PgcnPlan (the opaque plan handle standing in for the sparse tensor)
                          H = fp32 CUDA tensor

forward  = pack boundary rows -> all-to-all-v -> Z = A_local * [H_own ; H_halo]     (:123-127)
backward = G = A_local^T * gZ, halo-row partials sent back to their owners and SUMMED (:129-134;
           the reference ASSIGNS the received rows, quirk Q3 of SURVEY.md §8a — this op implements
           the intended semantics and the tests pin the difference).

Layouts (plan.layout):
  "local"  : H is [m, f] (owned rows only), Z is [m, f] — no n-sized tensors anywhere.
  "global" : H is [n, f] like the reference (rows it does not own are ignored = the reference's
             precondition Q0 that they are zero), Z is [n, f] with non-owned rows exactly 0.

All arithmetic happens in libpgcn_b200.so on the current CUDA stream; there is no CPU path.
"""
import torch

from . import cabi


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _check_feat(plan, H, rows, what):
    if not H.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the PGCN B200 path has no CPU fallback" % what)
    if H.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (what, H.dtype))
    if H.dim() != 2 or H.shape[0] != rows:
        raise ValueError("%s must be [%d, f], got %s" % (what, rows, tuple(H.shape)))
    if H.shape[1] > plan.f_max:
        raise ValueError("f=%d exceeds the plan's f_max=%d" % (H.shape[1], plan.f_max))
    return H.contiguous()


def aggregate_forward(plan, H_own, relu=False):
    """Z_own = (A * H)[owned rows]; H_own, Z_own are [m, f]. relu=True: Z_own = max(0, .), clamped inside the store of
    the launch that writes each row last (plan option "relu")."""
    H_own = _check_feat(plan, H_own, plan.m, "H")
    f = H_own.shape[1]
    Z = torch.empty((plan.m, f), dtype=torch.float32, device=H_own.device)
    lib = cabi.load()
    with torch.cuda.device(H_own.device):
        if relu:
            plan.set_option("relu", 1)
        try:
            cabi.check(lib.pgcn_forward(plan.handle, H_own.data_ptr(), Z.data_ptr(), f, _stream_ptr()), plan.handle)
        finally:
            if relu:
                plan.set_option("relu", 0)
    if plan.lp.k > 1:
        plan.count_exchange(backward=False)
    return Z


def aggregate_backward(plan, gZ_own):
    """G_own = (A^T * gZ)[owned rows] with every peer's contribution added; [m, f]."""
    gZ_own = _check_feat(plan, gZ_own, plan.m, "grad_output")
    f = gZ_own.shape[1]
    G = torch.empty((plan.m, f), dtype=torch.float32, device=gZ_own.device)
    lib = cabi.load()
    with torch.cuda.device(gZ_own.device):
        cabi.check(lib.pgcn_backward(plan.handle, gZ_own.data_ptr(), G.data_ptr(), f, _stream_ptr()), plan.handle)
    if plan.lp.k > 1:
        plan.count_exchange(backward=True)
    return G


class PSpMM(torch.autograd.Function):
    """Same call shape as the reference operator: PSpMM.apply(A, H) (GPU/PGCN.py:145)."""

    @staticmethod
    def forward(ctx, A, H):
        ctx.plan = A
        if A.layout == "global":
            _check_feat(A, H, A.n, "H")
            Z_own = aggregate_forward(A, H.index_select(0, A.owned_index()))
            Z = torch.zeros((A.n, H.shape[1]), dtype=torch.float32, device=H.device)
            Z.index_copy_(0, A.owned_index(), Z_own)
            return Z
        return aggregate_forward(A, H)

    @staticmethod
    def backward(ctx, grad_output):
        A = ctx.plan
        if A.layout == "global":
            g = _check_feat(A, grad_output, A.n, "grad_output")
            G_own = aggregate_backward(A, g.index_select(0, A.owned_index()))
            G = torch.zeros((A.n, g.shape[1]), dtype=torch.float32, device=g.device)
            G.index_copy_(0, A.owned_index(), G_own)
            return None, G
        return None, aggregate_backward(A, grad_output)


class PSpMMRelu(torch.autograd.Function):
    """relu(A * X) with the clamp fused into the aggregation's output store. Feeding it
    X = linear(H) gives the reference layer relu(linear(PSpMM(A, H))) of GPU/PGCN.py:144-148 up to fp32 association
    ((A H) W^T == A (H W^T)); the dense step then also runs before the aggregation instead of after it.
    Backward: relu's mask comes from the saved output (out > 0), then the regular PSpMM backward."""

    @staticmethod
    def forward(ctx, A, X):
        ctx.plan = A
        out = aggregate_forward(A, X, relu=True)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        return None, aggregate_backward(ctx.plan, grad_output * (out > 0))


# ---- the pieces, individually callable (NCCL transport), mirroring communicate_fgm ----------------

def spmm_local(plan, H_own, H_halo=None, transpose=False):
    """torch.sparse.mm(A, H) / torch.sparse.mm(A.t(), g) of GPU/PGCN.py:127,132 on the local matrix.
    transpose=False: returns Z [m, f].  transpose=True: returns (G_own [m, f], G_halo [h, f])."""
    lp = plan.lp
    H_own = _check_feat(plan, H_own, lp.m, "H_own")
    f = H_own.shape[1]
    dev = H_own.device
    lib = cabi.load()
    with torch.cuda.device(dev):
        if not transpose:
            if lp.h:
                H_halo = _check_feat(plan, H_halo, lp.h, "H_halo")
            Z = torch.empty((lp.m, f), dtype=torch.float32, device=dev)
            cabi.check(lib.pgcn_spmm(plan.handle, 0, H_own.data_ptr(), H_halo.data_ptr() if lp.h else None,
                                     Z.data_ptr(), None, f, _stream_ptr()), plan.handle)
            return Z
        G = torch.empty((lp.m, f), dtype=torch.float32, device=dev)
        Gh = torch.empty((lp.h, f), dtype=torch.float32, device=dev)
        cabi.check(lib.pgcn_spmm(plan.handle, 1, H_own.data_ptr(), None, G.data_ptr(),
                                 Gh.data_ptr() if lp.h else None, f, _stream_ptr()), plan.handle)
        return G, Gh


def spmm_split(plan, H_own, H_halo):
    """The overlapped forward's two kernels back to back: Z = A_own*H_own, then Z += A_halo*H_halo
    (Parallel-GCN/main.c:271,295). Same result as spmm_local up to fp32 summation order."""
    lp = plan.lp
    H_own = _check_feat(plan, H_own, lp.m, "H_own")
    H_halo = _check_feat(plan, H_halo, lp.h, "H_halo")
    f = H_own.shape[1]
    Z = torch.empty((lp.m, f), dtype=torch.float32, device=H_own.device)
    lib = cabi.load()
    with torch.cuda.device(H_own.device):
        cabi.check(lib.pgcn_spmm(plan.handle, 2, H_own.data_ptr(), None, Z.data_ptr(), None, f, _stream_ptr()), plan.handle)
        cabi.check(lib.pgcn_spmm(plan.handle, 3, None, H_halo.data_ptr(), Z.data_ptr(), None, f, _stream_ptr()), plan.handle)
    return Z


def pack_rows(plan, H_own):
    """send slab [S, f]: H[send_map[p]] for every peer p, concatenated in peer order (GPU/PGCN.py:104)."""
    H_own = _check_feat(plan, H_own, plan.lp.m, "H_own")
    f = H_own.shape[1]
    slab = torch.empty((plan.lp.S, f), dtype=torch.float32, device=H_own.device)
    with torch.cuda.device(H_own.device):
        cabi.check(cabi.load().pgcn_pack(plan.handle, H_own.data_ptr(), slab.data_ptr(), f, _stream_ptr()), plan.handle)
    return slab


def exchange_rows(plan, send_slab, backward=False):
    """The all-to-all-v of GPU/PGCN.py:99-115 (NCCL transport). Returns the received slab."""
    lp = plan.lp
    rows_out = lp.h if backward else lp.S
    rows_in = lp.S if backward else lp.h
    send_slab = _check_feat(plan, send_slab, rows_out, "send_slab")
    f = send_slab.shape[1]
    recv = torch.empty((rows_in, f), dtype=torch.float32, device=send_slab.device)
    with torch.cuda.device(send_slab.device):
        cabi.check(cabi.load().pgcn_exchange(plan.handle, send_slab.data_ptr(), recv.data_ptr(), f,
                                             1 if backward else 0, _stream_ptr()), plan.handle)
    plan.count_exchange(backward=backward)
    return recv


def unpack_add(plan, recv_slab, G_own):
    """G_own[send_idx[j]] += recv_slab[j] for all j, fixed order (in place). Returns G_own."""
    recv_slab = _check_feat(plan, recv_slab, plan.lp.S, "recv_slab")
    if not G_own.is_contiguous():
        # in/out argument: a silent .contiguous() copy would accumulate into a temporary and leave G_own unchanged
        raise ValueError("G_own is updated in place and must be contiguous")
    G_own = _check_feat(plan, G_own, plan.lp.m, "G_own")
    with torch.cuda.device(G_own.device):
        cabi.check(cabi.load().pgcn_unpack_add(plan.handle, recv_slab.data_ptr(), G_own.data_ptr(),
                                               G_own.shape[1], _stream_ptr()), plan.handle)
    return G_own


def communicate_fgm(plan, H, backward=False):
    """The exchange of GPU/PGCN.py:85-119 in the compact layout.
    forward : H [m, f] -> halo rows [h, f] (what the reference scatters into X[recv_map]).
    backward: halo partials [h, f] -> contributions for my rows, summed, as a dense [m, f]."""
    if not backward:
        return exchange_rows(plan, pack_rows(plan, H), backward=False)
    recv = exchange_rows(plan, H, backward=True)
    G = torch.zeros((plan.lp.m, H.shape[1]), dtype=torch.float32, device=H.device)
    return unpack_add(plan, recv, G)
