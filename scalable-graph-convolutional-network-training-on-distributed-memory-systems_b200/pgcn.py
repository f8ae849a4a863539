"""PGCN trainer CLI — the reference's surface (GPU/PGCN.py:136-286) over the B200 operator.

    python PGCN.py -a A.mtx -p A.mtx.8.hp -b nccl -s 8 -l 2 -f 128

Kept from the reference: flags -a -p -b -s -l -f (GPU/PGCN.py:262-278); rank/size from
SLURM_PROCID / SLURM_NPROCS (:258-260) with torchrun's RANK / WORLD_SIZE as a fallback; rendezvous
from MASTER_ADDR / MASTER_PORT; device cuda:{rank % ndev} (:169); inputs H[i, :] = i (:186-188) and
labels i % f (:192); L x (PSpMM -> Linear(f, f, bias=False) -> ReLU) (:136-148, :194-196); parameter
averaging at start (:156-160); Adam lr 1e-3 (:200); 1 warm-up + 4 timed epochs (:202-226); gradient
all-reduce / world_size (:150-154); stdout fields `Epoch {:05d} | Loss {:.4f}`, the per-rank stats
dict, `Elapsed time {:.4f}`, `total_vol: .. total_nmsg: ..` (:224-238).

Different by design: every tensor is [m_local, f] instead of [n, f]; the loss is
the reference's value computed from owned rows only, loss = (sum_owned nll + (n - m) * log f) / n,
which has exactly the reference's gradients (its non-owned rows are all-zero logits); the elapsed
time is bracketed by torch.cuda.synchronize(); `-b gloo` is refused — the B200 path has no CPU
fallback (the CPU oracle lives under oracle/ and is test infrastructure).
`--ref-quirks` reproduces quirk Q1 (halo rows counted twice in layer 1 because `run` feeds a fully
populated H into `H + X`, GPU/PGCN.py:117,186) so loss curves can be compared with seeded weights.
"""
import getopt
import math
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import graphio, plan as planmod
from .op import PSpMM, PSpMMRelu, communicate_fgm, spmm_local, aggregate_backward


class _PSpMMQuirkQ1(torch.autograd.Function):
    """Layer-1 aggregation as the reference computes it when fed an unmasked H: A_loc (H + 1_halo H),
    i.e. halo rows weigh twice (quirk Q1). Backward is the regular one."""

    @staticmethod
    def forward(ctx, A, H):
        ctx.plan = A
        halo = communicate_fgm(A, H, backward=False)
        return spmm_local(A, H, 2.0 * halo if A.lp.h else None)

    @staticmethod
    def backward(ctx, g):
        return None, aggregate_backward(ctx.plan, g)


class PGCN(nn.Module):
    """GPU/PGCN.py:136-148 with the plan handle in place of the sparse tensor."""

    def __init__(self, A, in_features, out_features, quirk_q1=False, fused=False):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=False)
        self.A = A
        self.quirk_q1 = quirk_q1
        self.fused = fused and not quirk_q1

    def forward(self, H):
        if self.fused:
            # relu(A (H W^T)): dense step on the m owned rows first, relu fused into the aggregation's store
            return PSpMMRelu.apply(self.A, self.linear(H))
        H = _PSpMMQuirkQ1.apply(self.A, H) if self.quirk_q1 else PSpMM.apply(self.A, H)
        H = self.linear(H)
        return F.relu(H)


def average_gradients(model, world_size):
    for param in model.parameters():                      # GPU/PGCN.py:150-154
        dist.all_reduce(param.grad.data, op=dist.ReduceOp.SUM)
        param.grad.data /= world_size


def initialize_parameters(model, world_size):
    for param in model.parameters():                      # GPU/PGCN.py:156-160
        dist.all_reduce(param.data, op=dist.ReduceOp.SUM)
        param.data /= world_size


def reference_loss(logits_own, labels_own, n):
    """F.nll_loss(log_softmax(logits), labels) over ALL n rows as the reference computes it
    (GPU/PGCN.py:204-205), from the owned rows: non-owned rows are all-zero logits -> nll = log f."""
    m, f = logits_own.shape
    nll = F.nll_loss(F.log_softmax(logits_own, 1), labels_own, reduction="sum") if m else logits_own.sum()
    return (nll + (n - m) * math.log(f)) / n


def run(rank, size, nlayers, nfeatures, path_A, path_partvec, backend, ref_quirks=False, transport="auto",
        out=sys.stdout, seed=None, fused=False):
    if backend != "nccl":
        raise RuntimeError("backend '%s': the B200 PGCN path runs on CUDA devices over NCCL/NVLink only "
                           "(no CPU fallback); use -b nccl" % backend)
    device = torch.device("cuda", rank % torch.cuda.device_count())      # GPU/PGCN.py:169
    torch.cuda.set_device(device)
    cache = os.environ.get("PGCN_PLAN_CACHE")
    if cache:                                                            # optional on-disk plan cache (§8f rank 2)
        lp_host = planmod.cached_local_plan(path_A, path_partvec, rank, size, cache)
    else:
        A = graphio.read_adjacency(path_A)                               # :171
        partvec = graphio.read_partvec(path_partvec, A.shape[0])         # :172-173
        graphio.check_partvec(partvec, size)
        lp_host = planmod.build_local_plan(A, partvec, rank, size)       # :175-176
    n = lp_host.n
    plan = planmod.PgcnPlan(lp_host, nfeatures, device=device)           # :178-182
    if ref_quirks:
        transport = "nccl"            # the quirk emulation drives the exchange step by step (NCCL entry points)
    used = plan.init_comm(transport=transport)
    plan.autotune(nfeatures)
    lp = plan.lp

    own = torch.from_numpy(lp.owned).to(device)
    H = own.to(torch.float32).unsqueeze(1).repeat(1, nfeatures).contiguous().requires_grad_(True)   # :186-188
    labels = own % nfeatures                                                                          # :192

    if seed is not None:
        torch.manual_seed(seed)
    model = nn.Sequential(*[PGCN(plan, nfeatures, nfeatures, quirk_q1=(ref_quirks and i == 0), fused=fused)
                            for i in range(nlayers)]).to(device)        # :194-198
    if size > 1:
        initialize_parameters(model, size)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)            # :200

    def epoch():
        logits = model(H)
        loss = reference_loss(logits, labels, n)
        optimizer.zero_grad()
        loss.backward()
        if size > 1:
            average_gradients(model, size)
        optimizer.step()
        return loss

    epoch()                                                              # warm-up epoch, :202-209
    torch.cuda.synchronize()
    start = time.time()
    losses = []
    for ep in range(4):                                                  # :212-224
        loss = epoch()
        losses.append(float(loss))
        if rank == 0:
            print("Epoch {:05d} | Loss {:.4f}".format(ep, losses[-1]), file=out, flush=True)
    torch.cuda.synchronize()
    elapsed = torch.tensor([time.time() - start], device=device)
    if size > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)                   # :228

    print(plan.stats, file=out, flush=True)                              # :230
    tot = torch.tensor([plan.stats["send_volume"], plan.stats["send_nmsg"]], device=device)
    if size > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)                       # :233-234
    if rank == 0:
        print("Elapsed time {:.4f}".format(elapsed.item()), file=out, flush=True)
        print("total_vol: {} total_nmsg: {}".format(int(tot[0]), int(tot[1])), file=out, flush=True)
    result = {"losses": losses, "total_vol": int(tot[0]), "total_nmsg": int(tot[1]), "elapsed": float(elapsed.item()),
              "transport": used}
    plan.close()
    return result


def init_process(rank, size, fn, nlayers, nfeatures, path_A, path_partvec, backend, **kw):
    if backend == "nccl" and torch.cuda.is_available():
        torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend, rank=rank, world_size=size)         # GPU/PGCN.py:242
    env = {k: os.environ.get(k) for k in ("MASTER_ADDR", "MASTER_PORT", "RANK", "WORLD_SIZE")}
    print("[{}] Initializing process group with: {}".format(os.getpid(), env), flush=True)
    try:
        return fn(rank, size, nlayers, nfeatures, path_A, path_partvec, backend, **kw)
    finally:
        dist.destroy_process_group()


def main(argv):
    size = int(os.environ.get("SLURM_NPROCS", os.environ.get("WORLD_SIZE", "1")))      # GPU/PGCN.py:258-260
    rank = int(os.environ.get("SLURM_PROCID", os.environ.get("RANK", "0")))
    os.environ["RANK"] = str(rank)
    try:
        opts, _ = getopt.getopt(argv, "a:p:b:s:l:f:", ["ref-quirks", "transport=", "seed=", "fused"])
    except getopt.GetoptError:
        print("a:p:b:", flush=True)                                       # the reference's usage text, :264
        sys.exit(2)
    path_A = path_partvec = None
    backend = "nccl"
    nlayers = nfeatures = None
    kw = {}
    for opt, arg in opts:
        if opt == "-a":
            path_A = arg
        elif opt == "-p":
            path_partvec = arg
        elif opt == "-b":
            backend = arg
        elif opt == "-s":
            size = int(arg)
        elif opt == "-l":
            nlayers = int(arg)
        elif opt == "-f":
            nfeatures = int(arg)
        elif opt == "--ref-quirks":
            kw["ref_quirks"] = True
        elif opt == "--transport":
            kw["transport"] = arg
        elif opt == "--seed":
            kw["seed"] = int(arg)
        elif opt == "--fused":
            kw["fused"] = True           # relu(A (H W^T)) with the clamp fused into the aggregation
    if path_A is None or path_partvec is None or nlayers is None or nfeatures is None:
        print("usage: PGCN.py -a <A.mtx> -p <partvec> -b nccl -s <nparts> -l <nlayers> -f <nfeatures>", flush=True)
        sys.exit(2)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    os.environ["WORLD_SIZE"] = str(size)
    init_process(rank, size, run, nlayers, nfeatures, path_A, path_partvec, backend, **kw)


if __name__ == "__main__":
    main(sys.argv[1:])
