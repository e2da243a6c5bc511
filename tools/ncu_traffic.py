#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (exported with `ncu -i X.ncu-rep --page raw --csv > raw.csv`) into
profiles/traffic.json: per kernel, the DRAM bytes one launch moved and the hash of the kernel sources the
capture was taken from.  bench.py reports `roofline.traffic` from this file only while the hash still matches.

  python tools/ncu_traffic.py profiles/r2_gru_step_full_raw.csv --kernel gru_step_tc_kernel --rows 32768
"""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * mult[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--rows", type=int, required=True)
    ap.add_argument("--launch", type=int, default=-1, help="which captured launch of the kernel to use (default: last)")
    ap.add_argument("--csrc-hash", default=None, help="hash of the sources the capture ran (default: the tree's current hash)")
    a = ap.parse_args()
    from bench import csrc_hash
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    name_col = col.get("Kernel Name", col.get("Function Name"))
    cand = [r for r in rows[2:] if len(r) == len(hdr) and a.kernel in r[name_col]]
    if not cand:
        raise SystemExit(f"no launch of {a.kernel} in {a.raw_csv}")
    r = cand[a.launch]
    g = lambda m: (r[col[m]], units[col[m]])
    rec = {"rows": a.rows, "dram_bytes_read": int(to_bytes(*g("dram__bytes_read.sum"))), "dram_bytes_write": int(to_bytes(*g("dram__bytes_write.sum"))),
           "gpu_time_us_under_ncu": float(g("gpu__time_duration.sum")[0].replace(",", "")) * {"us": 1, "usecond": 1, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(g("gpu__time_duration.sum")[1], 1),
           "sm__pipe_tensor_cycles_active_pct": float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")[0]),
           "source": os.path.relpath(os.path.abspath(a.raw_csv), ROOT), "csrc_hash": a.csrc_hash or csrc_hash(), "launches_in_capture": len(cand)}
    out = os.path.join(ROOT, "profiles", "traffic.json")
    db = json.load(open(out)) if os.path.exists(out) else {}
    db[a.kernel] = rec
    json.dump(db, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
