"""The peer-memory transport of the halo exchange on ONE GPU (so the driver's single-GPU box sees it run):

  * every rank's plan in this process on cuda:0, wired with plan.link_local_plans (same-process peers are reached
    through plain device pointers): fused put + epoch-signal kernels, per-peer wait kernels, the per-peer pipelined
    forward (own columns while rows travel, each source's block as it lands) and backward (partials for each peer
    first, leaving while the rest is computed) — against the fp64 truth, the reference's golden outputs, the
    non-overlapped path, and repeated calls (epoch parity of the double-buffered slabs);
  * two PROCESSES sharing cuda:0 through CUDA IPC and PgcnPlan.init_comm(transport="p2p") over gloo.

NCCL refuses two ranks on one device; its path is covered by tests/test_multigpu.py on >= 2 GPUs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import Golden, assert_close_fp32, fp32_tol
from oracle import pgcn_oracle as orc
from pgcn_b200 import cabi, graphio, plan as planmod

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev():
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: -m gpu tests must run on the B200 box")
    return torch.device("cuda", 0)


def run_all(plans, fn_name, inputs, f):
    """Enqueue one fused forward/backward per rank, each on its own stream (nothing blocks on the host: the
    ranks' kernels wait for each other on the device), then synchronise."""
    lib = cabi.load()
    outs = [torch.empty((p.lp.m, f), device=dev()) for p in plans]
    streams = [torch.cuda.Stream(device=dev()) for _ in plans]
    torch.cuda.synchronize()
    for p, x, y, s in zip(plans, inputs, outs, streams):
        cabi.check(getattr(lib, fn_name)(p.handle, x.data_ptr(), y.data_ptr(), f, s.cuda_stream), p.handle)
    torch.cuda.synchronize()
    return outs


@pytest.mark.parametrize("case", ["gemat11_k2", "gemat11_k3_hp", "gemat11_k3_rp", "karate_k3_hp", "rmat_k4"])
def test_peer_transport_all_ranks_on_one_gpu(case):
    if case == "rmat_k4":
        n, f, k = 12000, 128, 4
        A = graphio.synthetic_graph(n, 240000, seed=4)
        pv = graphio.random_partvec(n, k, seed=9)
        rs = np.random.RandomState(3)
        H = rs.uniform(-1, 1, size=(n, f)).astype(np.float32)
        G = rs.uniform(-1, 1, size=(n, f)).astype(np.float32)
        g = None
    else:
        g = Golden(case)
        A, pv, H, G, f, k = g.A, g.partvec, g.H, g.G, g.f, g.k
    plans = [planmod.build_plan(A, pv, r, k, f, device=dev()) for r in range(k)]
    assert planmod.link_local_plans(plans) == "p2p"
    Z64 = orc.truth_forward(A, H); G64 = orc.truth_backward(A, G)
    tolZ = fp32_tol(A, H, int(orc.row_degree(A).max())); tolG = fp32_tol(A.T, G, int(orc.row_degree(A.T).max()))
    Hd = [torch.from_numpy(H[p.lp.owned]).to(dev()) for p in plans]
    Gd = [torch.from_numpy(G[p.lp.owned]).to(dev()) for p in plans]
    res = {}
    for overlap in (1, 0):
        for p in plans:
            assert p.get_option("p2p") == 1
            p.set_option("overlap", overlap)
        l0 = sum(p.launch_count() for p in plans)
        Z = run_all(plans, "pgcn_forward", Hd, f)
        Gr = run_all(plans, "pgcn_backward", Gd, f)
        assert sum(p.launch_count() for p in plans) - l0 >= 2 * k * (k - 1)      # put + wait kernels ran
        for r, p in enumerate(plans):
            own = p.lp.owned
            assert_close_fp32(Z[r].cpu().numpy(), Z64[own], tolZ[own], "%s fwd r%d overlap=%d" % (case, r, overlap))
            assert_close_fp32(Gr[r].cpu().numpy(), G64[own], tolG[own], "%s bwd r%d overlap=%d" % (case, r, overlap))
            if g is not None:
                np.testing.assert_allclose(Z[r].cpu().numpy(), g.get(r, "Z1_own"), rtol=2e-5,
                                           atol=2e-6 * max(1.0, np.abs(Z64).max()))
        res[overlap] = (Z, Gr)
    for r in range(k):
        torch.testing.assert_close(res[1][0][r], res[0][0][r], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(res[1][1][r], res[0][1][r], rtol=1e-4, atol=1e-5)
    # repeated exchanges: slabs are re-used every second epoch, results must not change by a bit
    for _ in range(4):
        Z = run_all(plans, "pgcn_forward", Hd, f)
    for r in range(k):
        assert torch.equal(Z[r], res[0][0][r])
    for p in plans:
        p.close()


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import Golden, fp32_tol
from oracle import pgcn_oracle as orc
from pgcn_b200 import plan as planmod
from pgcn_b200.op import PSpMM
rank, k = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=k)
g = Golden(%(case)r)
d = torch.device("cuda", 0)
torch.cuda.set_device(d)
p = planmod.build_plan(g.A, g.partvec, rank, k, g.f, device=d)
used = p.init_comm(transport="p2p", nccl_fallback=False)
assert used == "p2p", used
own = p.lp.owned
Hd = torch.from_numpy(g.H[own]).to(d).requires_grad_(True)
Z = PSpMM.apply(p, Hd)
Z.backward(torch.from_numpy(g.G[own]).to(d))
for _ in range(3):
    Z2 = PSpMM.apply(p, Hd.detach())
torch.cuda.synchronize()
assert torch.equal(Z2, Z.detach())
Z64 = orc.truth_forward(g.A, g.H)[own]; G64 = orc.truth_backward(g.A, g.G)[own]
tz = fp32_tol(g.A, g.H, int(orc.row_degree(g.A).max()))[own]
tg = fp32_tol(g.A.T, g.G, int(orc.row_degree(g.A.T).max()))[own]
assert (np.abs(Z.detach().cpu().numpy() - Z64) <= tz).all()
assert (np.abs(Hd.grad.cpu().numpy() - G64) <= tg).all()
dist.barrier()
p.close()
print("rank %%d ok" %% rank)
"""


def test_peer_transport_two_processes_one_gpu_over_cuda_ipc(tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    case, k = "gemat11_k2", 2
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "case": case})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE=str(k))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(k)]
    outs = []
    for pr in procs:
        try:
            out, _ = pr.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("two-process peer-transport worker timed out")
        outs.append(out)
    for r, (pr, out) in enumerate(zip(procs, outs)):
        assert pr.returncode == 0, "rank %d failed:\n%s" % (r, out[-3000:])
        assert "rank %d ok" % r in out


def test_more_ranks_than_the_peer_transport_takes():
    """k = 17 > 16: plans must build and the step-by-step pieces (pack -> wire order -> SpMM, A^T g -> scatter-add)
    must work — the NCCL transport has no rank limit; only pgcn_p2p_export refuses (fixed-size peer arrays)."""
    import ctypes as C
    from test_gpu_parity import forward_all, backward_all
    n, f, k = 1700, 32, 17
    A = graphio.synthetic_graph(n, 20000, seed=8)
    pv = graphio.random_partvec(n, k, seed=2)
    H = np.random.RandomState(1).uniform(-1, 1, size=(n, f)).astype(np.float32)
    plans = [planmod.build_plan(A, pv, r, k, f, device=dev()) for r in range(k)]
    Z = forward_all(plans, H)
    Gd = backward_all(plans, H)
    Z64 = orc.truth_forward(A, H); G64 = orc.truth_backward(A, H)
    tolZ = fp32_tol(A, H, int(orc.row_degree(A).max())); tolG = fp32_tol(A.T, H, int(orc.row_degree(A.T).max()))
    for r, p in enumerate(plans):
        own = p.lp.owned
        assert_close_fp32(Z[r].cpu().numpy(), Z64[own], tolZ[own], "k=17 fwd r%d" % r)
        assert_close_fp32(Gd[r].cpu().numpy(), G64[own], tolG[own], "k=17 bwd r%d" % r)
    blob = C.create_string_buffer(cabi.P2P_HANDLE_BYTES)
    assert cabi.load().pgcn_p2p_export(plans[0].handle, blob) < 0          # refused, with a message
    assert b"16" in cabi.load().pgcn_last_error(plans[0].handle)
    for p in plans:
        p.close()
