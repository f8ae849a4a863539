"""Golden loss curve of the UNMODIFIED mini-batch trainer GPU/PGCN-Mini-batch.py (its own run(), gloo, one rank — no
peers, so none of the exchange quirks Q1-Q3 can bite) on the shipped karate graph: 3 layers, f = 4, batch_size = 12,
weights seeded with torch.manual_seed(1234) right before run() builds the model.
    python tests/golden/make_minibatch_e2e_golden.py     ->  tests/golden/karate_minibatch_e2e.json"""
import contextlib
import importlib.util
import io
import json
import os
import pickle
import re
import sys
import tempfile
import warnings

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    warnings.filterwarnings("ignore")
    spec = importlib.util.spec_from_file_location("ref_mb", os.path.join(REF, "GPU", "PGCN-Mini-batch.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", RANK="0", WORLD_SIZE="1")
    dist.init_process_group("gloo", rank=0, world_size=1)
    n = 34
    with tempfile.TemporaryDirectory() as d:
        pv = os.path.join(d, "pv1.pkl")
        pickle.dump([0] * n, open(pv, "wb"))
        buf = io.StringIO()
        torch.manual_seed(1234)
        with contextlib.redirect_stdout(buf):
            ref.run(0, 1, 3, 4, os.path.join(REF, "GPU/SHP/data/karate/karate.mtx"), pv, "gloo", 12)
    dist.destroy_process_group()
    text = buf.getvalue()
    losses = [float(x) for x in re.findall(r"Loss ([0-9.eE+-]+)", text)]
    vol = re.search(r"total_vol: (\d+) total_nmsg: (\d+)", text)
    out = {"graph": "GPU/SHP/data/karate/karate.mtx", "k": 1, "layers": 3, "f": 4, "batch_size": 12, "seed": 1234,
           "losses": losses, "total_vol": int(vol.group(1)), "total_nmsg": int(vol.group(2)), "stdout": text}
    json.dump(out, open(os.path.join(HERE, "karate_minibatch_e2e.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "stdout"}))


if __name__ == "__main__":
    sys.exit(main())
