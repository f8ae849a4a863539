"""The shipped part vectors (bench_data/, produced by tools/make_partvecs.py with the reference's PaToH / METIS libraries)
and the way bench.py picks its workload: host-only checks."""
import argparse
import glob
import os
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)


def test_shipped_part_vectors_are_well_formed():
    from pgcn_b200 import graphio
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", "*.npz")))
    assert any(os.path.basename(f) == "C5.8.hp.npz" for f in files), "the scaling config's hp vector must ship"
    for path in files:
        cfg, k, method = os.path.basename(path)[:-4].split(".")
        k = int(k)
        z = np.load(path)
        pv = z["partvec"]
        assert pv.dtype == np.uint8 and pv.shape[0] == graphio.CONFIGS[cfg][0]
        counts = np.bincount(pv, minlength=k)
        assert counts.shape[0] == k and (counts > 0).all(), path
        assert method in ("hp", "gp") and str(z["tool"]) == "oracle/part_tool.cpp"
        assert ("cut" in str(z["report"])) and "halo rows total" in str(z["report"])


def test_bench_picks_config_and_partition_like_the_docs_say():
    import bench
    old = sys.argv
    try:
        sys.argv = ["bench.py"]
        a1 = bench.parse()
        sys.argv = ["bench.py", "--gpus", "8"]
        a8 = bench.parse()
        sys.argv = ["bench.py", "--gpus", "8", "--impl", "reference"]
        r8 = bench.parse()
    finally:
        sys.argv = old
    assert a1.config == "C2" and a8.config == "C5" and r8.config == "C5"
    # both arms name the workload with the same string (the driver compares them)
    assert bench.workload_name(a8.config) == bench.workload_name(r8.config)
    from pgcn_b200 import graphio
    n = graphio.CONFIGS["C5"][0]
    pv, name = bench.part_vector(a8, n, 8)
    assert name.startswith("hp") and pv.shape[0] == n and pv.max() == 7
    a8.partition = "rp"
    assert bench.part_vector(a8, n, 8)[1].startswith("rp")
    assert bench.part_vector(a1, graphio.CONFIGS["C2"][0], 1)[1] == "single part"
    assert len(bench.source_hash()) == 16


def test_shipped_vector_belongs_to_the_generated_graph(tmp_path):
    """The part vectors are only meaningful for the exact graph graphio generates: recompute the halo volume of
    C2 / k = 8 / hp from a freshly generated C2 and compare with the number the partitioner tool reported when the
    vector was made (any change of the generator's output would break this)."""
    import re
    from pgcn_b200 import graphio
    z = np.load(os.path.join(ROOT, "bench_data", "C2.8.hp.npz"))
    reported = int(re.search(r"halo rows total=(\d+)", str(z["report"])).group(1))
    A = graphio.config_graph("C2", cache_dir=None).tocoo()
    pv = z["partvec"].astype(np.int64)
    prow, pcol = pv[A.row], pv[A.col]
    cross = prow != pcol
    keys = np.unique(prow[cross] * A.shape[0] + A.col[cross])
    assert keys.shape[0] == reported
