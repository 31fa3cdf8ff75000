#!/usr/bin/env python3
"""Turn the raw export of an `ncu --set full` capture into the tracked figures bench.py reads.

    ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > profiles/r2_ncu_raw.csv
    python scripts/ncu_to_traffic.py profiles/r2_ncu_raw.csv --matrices 296 --order 1500 --capture "r2, prof_target.py 296 500"

Writes profiles/ncu_traffic.json: per kernel the DRAM bytes read / written by one launch, the
number of matrices in that launch and the sha256 of the kernel's source file at capture time
(bench.py reports `traffic: null` when the source has changed since).
"""
import argparse
import csv
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = {"mf_potrf_ll_kernel": "miniframe_b200/csrc/mf_potrf.cu",
           "mf_cov_build_big_il_kernel": "miniframe_b200/csrc/mf_cov.cu",
           "mf_cov_build_big_lower_kernel": "miniframe_b200/csrc/mf_cov.cu"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--matrices", type=int, required=True)
    ap.add_argument("--order", type=int, required=True)
    ap.add_argument("--capture", required=True)
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        key = next((k for k in SOURCES if k in name), None)
        if key is None:
            continue
        val = {}
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            val[m] = float(r[ix[m]].replace(",", "")) * UNIT[units[ix[m]]]
        src = SOURCES[key]
        table[key] = {"kernel": name, "matrices": args.matrices, "matrix_order": args.order,
                      "dram_bytes_read": val["dram__bytes_read.sum"], "dram_bytes_write": val["dram__bytes_write.sum"],
                      "duration_ms": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) *
                                     {"msecond": 1.0, "ms": 1.0, "usecond": 1e-3, "us": 1e-3, "second": 1e3, "s": 1e3, "nsecond": 1e-6, "ns": 1e-6}[units[ix["gpu__time_duration.sum"]]],
                      "source": src, "source_sha256": hashlib.sha256(open(os.path.join(ROOT, src), "rb").read()).hexdigest(),
                      "capture": args.capture, "raw_export": os.path.relpath(args.csv, ROOT)}
        print(key, json.dumps(table[key])[:300])
    json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
