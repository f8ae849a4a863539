"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/GPU/PGCN.py).

Run in the build container (the reference is mounted read-only there; it does not exist on the GPU
box, which is why the vectors are committed):

    python tests/golden/make_golden.py

For every case (fixture matrix x part vector x k) one gloo process per rank imports the reference
module with importlib (its `__main__` guard keeps it from running), sets the module globals its
`run` would set (GPU/PGCN.py:163-189), and records

  * send_map / recv_map from the reference's compute_communication_maps        (GPU/PGCN.py:37-51)
  * the rows kept by get_partitiont_of_adjacency_matrix                         (GPU/PGCN.py:53-64)
  * Z1 = PSpMM.apply(A, H) under the Q0 precondition (H zero on non-owned rows, X fresh zeros)
  * Hgrad = the gradient PSpMM.backward returns for a seeded upstream gradient   (literal: Q2, Q3)
  * Z2 = a second forward on the same H with the now-stale X                     (literal: Q2)
  * the stats counters after those three exchanges                               (GPU/PGCN.py:78-83)

and, once, an end-to-end run of the reference's own `run` on gemat11 / 3 ranks for the stdout
fields (total_vol, total_nmsg, seeded losses).

The fixture matrices are public SuiteSparse data shipped inside the reference repo; they are stored
in the .npz as COO arrays so the tests do not need /root/reference.
"""
import importlib.util
import io
import json
import os
import pickle
import sys
from contextlib import redirect_stdout

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.io import mmread

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SEED = 20260922


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_pgcn", os.path.join(REF, "GPU", "PGCN.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def case_inputs(n, f, partvec, rank):
    """Seeded H (n x f) and upstream gradient; each rank sees H only on the rows it owns (Q0)."""
    rng = np.random.RandomState(SEED)
    H = rng.uniform(-1.0, 1.0, size=(n, f)).astype(np.float32)
    G = rng.uniform(-1.0, 1.0, size=(n, f)).astype(np.float32)
    own = np.asarray(partvec) == rank
    Hr = H.copy()
    Hr[~own] = 0.0
    return H, G, Hr


def worker(rank, size, port, mtx_path, partvec, f, out_dir, tag):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    import warnings
    warnings.filterwarnings("ignore")
    ref = load_reference()
    ref.myrank, ref.world_size, ref.device = rank, size, torch.device("cpu")
    A = mmread(mtx_path).tocoo()
    n = A.shape[0]
    ref.send_map, ref.recv_map = ref.compute_communication_maps(A, partvec, rank, size)
    Aloc = ref.get_partitiont_of_adjacency_matrix(A, partvec, rank)
    ref.send_buffers, ref.recv_buffers = {}, {}
    for src, idx in ref.recv_map.items():
        ref.recv_buffers[src] = torch.zeros(len(idx), f)
    for tgt, idx in ref.send_map.items():
        ref.send_buffers[tgt] = torch.zeros(len(idx), f)
    ref.init_stats()
    H, G, Hr = case_inputs(n, f, partvec, rank)
    Ht = torch.tensor(Hr, requires_grad=True)
    ref.X = torch.zeros(Ht.shape)
    Z1 = ref.PSpMM.apply(Aloc, Ht)
    Z1.backward(torch.tensor(G))
    Hgrad = Ht.grad.detach().clone()
    with torch.no_grad():
        Z2 = ref.PSpMM.apply(Aloc, Ht.detach())
    Ac = Aloc.coalesce()
    rec = {
        "Z1": Z1.detach().numpy(), "Hgrad": Hgrad.numpy(), "Z2": Z2.numpy(),
        "loc_row": Aloc._indices()[0].numpy(), "loc_col": Aloc._indices()[1].numpy(),
        "loc_val": Aloc._values().numpy(), "loc_nnz_coalesced": np.array(Ac._nnz()),
        "stats": np.array([int(ref.stats[k]) for k in ("send_volume", "recv_volume", "send_nmsg", "recv_nmsg")]),
    }
    for p, idx in ref.send_map.items():
        rec["send_%d" % p] = idx.numpy()
    for p, idx in ref.recv_map.items():
        rec["recv_%d" % p] = idx.numpy()
    np.savez(os.path.join(out_dir, "_tmp_%s_r%d.npz" % (tag, rank)), **rec)
    dist.barrier()
    dist.destroy_process_group()


def e2e_worker(rank, size, port, mtx_path, pv_path, nlayers, f, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(size)
    import warnings
    warnings.filterwarnings("ignore")
    ref = load_reference()
    torch.manual_seed(1000 + rank)
    buf = io.StringIO()
    with redirect_stdout(buf):
        ref.init_process(rank, size, ref.run, nlayers, f, mtx_path, pv_path, "gloo")
    with open(os.path.join(out_dir, "_tmp_e2e_r%d.txt" % rank), "w") as fh:
        fh.write(buf.getvalue())


def run_case(tag, mtx_path, partvec, k, f, port):
    partvec = [int(p) for p in partvec]
    mp.spawn(worker, args=(k, port, mtx_path, partvec, f, HERE, tag), nprocs=k, join=True)
    A = mmread(mtx_path).tocoo()
    out = {"n": np.array(A.shape[0]), "k": np.array(k), "f": np.array(f), "seed": np.array(SEED),
           "row": A.row.astype(np.int32), "col": A.col.astype(np.int32), "val": A.data.astype(np.float64),
           "partvec": np.array(partvec, dtype=np.int32)}
    pv = np.array(partvec)
    for r in range(k):
        path = os.path.join(HERE, "_tmp_%s_r%d.npz" % (tag, r))
        z = np.load(path)
        own = np.flatnonzero(pv == r)
        nz_rows = np.flatnonzero(np.abs(z["Z1"]).sum(axis=1) != 0)
        assert np.all(np.isin(nz_rows, own)), "reference Z has non-zero non-owned rows"
        out["r%d_Z1_own" % r] = z["Z1"][own]
        out["r%d_Z2_own" % r] = z["Z2"][own]
        out["r%d_Hgrad_own" % r] = z["Hgrad"][own]
        for key in z.files:
            if key.startswith(("send_", "recv_", "loc_", "stats")):
                out["r%d_%s" % (r, key)] = z[key]
        os.remove(path)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag, {k_: v.shape for k_, v in out.items() if k_.endswith("Z1_own")})


def main():
    gem = os.path.join(REF, "GPU/hypergraph/data/gemat11/gemat11.mtx")
    hp3 = list(map(int, open(os.path.join(REF, "GPU/hypergraph/data/gemat11.mtx.3.hp")).readline().split()))
    rp3 = list(map(int, open(os.path.join(REF, "GPU/hypergraph/data/gemat11.mtx.3.rp")).readline().split()))
    kar = os.path.join(REF, "GPU/SHP/data/karate/karate.mtx")
    khp = pickle.load(open(os.path.join(REF, "GPU/SHP/data/partvec.hp.3"), "rb"))
    kst = pickle.load(open(os.path.join(REF, "GPU/SHP/data/partvec.stchp.3"), "rb"))
    port = 29710
    cases = [
        ("gemat11_k1", gem, [0] * len(hp3), 1, 16),
        ("gemat11_k2", gem, [min(p, 1) for p in hp3], 2, 16),
        ("gemat11_k3_hp", gem, hp3, 3, 16),
        ("gemat11_k3_rp", gem, rp3, 3, 16),
        ("karate_k3_hp", kar, khp, 3, 8),
        ("karate_k3_stchp", kar, kst, 3, 20),
    ]
    for tag, path, pv, k, f in cases:
        if "--e2e-only" not in sys.argv:
            run_case(tag, path, pv, k, f, port)
        port += 1

    # end-to-end stdout of the reference's own run(): gemat11, -l 2 -f 16, gloo, k = 1, 2, 3
    # (k = 3 uses the shipped .3.hp; k = 2 merges its parts 1 and 2; k = 1 is all zeros)
    import tempfile
    from pgcn_b200 import graphio
    L, f = 2, 16
    allrec = {}
    for k, pv in ((1, [0] * len(hp3)), (2, [min(p, 1) for p in hp3]), (3, hp3)):
        port += 1
        with tempfile.TemporaryDirectory() as td:
            pv_path = os.path.join(td, "gemat11.mtx.%d.hp" % k)
            graphio.write_partvec(pv_path, pv)
            mp.spawn(e2e_worker, args=(k, port, gem, pv_path, L, f, HERE), nprocs=k, join=True)
        txt = open(os.path.join(HERE, "_tmp_e2e_r0.txt")).read()
        for r in range(k):
            os.remove(os.path.join(HERE, "_tmp_e2e_r%d.txt" % r))
        rec = {"cmd": "PGCN.py -a gemat11.mtx -p <partvec k=%d> -b gloo -s %d -l 2 -f 16 (torch.manual_seed(1000+rank))" % (k, k),
               "losses": [], "total_vol": None, "total_nmsg": None}
        for line in txt.splitlines():
            if line.startswith("Epoch"):
                rec["losses"].append(float(line.split("Loss")[1]))
            if line.startswith("total_vol"):
                parts = line.replace(":", " ").split()
                rec["total_vol"], rec["total_nmsg"] = int(parts[1]), int(parts[3])
        allrec["k%d" % k] = rec
        print(k, rec)
    json.dump(allrec, open(os.path.join(HERE, "gemat11_e2e.json"), "w"), indent=1)
    json.dump(allrec["k3"], open(os.path.join(HERE, "gemat11_k3_hp_e2e.json"), "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
