#!/usr/bin/env python
"""Turn an ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) into
profiles/traffic_<config>.json, stamped with the hash of the kernel sources it was measured on (bench.py quotes the
number as roofline.traffic only while that hash matches).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:spmm_ring -s 4 -c 3 --csv --log-file gpurun_out/traffic_raw.csv python tools/tune_spmm.py --config C2 --single kernel=0 --iters 3
    python tools/update_traffic.py gpurun_out/traffic_raw.csv C2 [out.json]
"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    path, config = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic_%s.json" % config)
    rows = [r for r in csv.reader(open(path)) if r]
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    per = {}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        d = per.setdefault(r[iid], {"kernel": r[ik]})
        d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    launches = [d for d in per.values() if "dram__bytes_read.sum" in d]
    rd = sum(d["dram__bytes_read.sum"] for d in launches) / len(launches)
    wr = sum(d["dram__bytes_write.sum"] for d in launches) / len(launches)
    us = sum(d.get("gpu__time_duration.sum", 0.0) for d in launches) / len(launches)
    import bench
    rec = {"config": config, "kernel": launches[0]["kernel"][:80], "launches_averaged": len(launches),
           "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
           "ncu_us_per_launch": us, "source_hash": bench.source_hash(),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, tools/tune_spmm.py --single (default options)"}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
