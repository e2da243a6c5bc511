#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`).
    python tools/launch_summary.py gpurun_out/launches.csv [--skip N] [--take N] > profiles/rN_launches_summary.txt
--skip / --take select a window of launches (e.g. the one timed step after the warm-up step)."""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=0)
    ap.add_argument("--take", type=int, default=0)
    ap.add_argument("--title", default="")
    ap.add_argument("--marker", default="", help="kernel that ends a step (e.g. adam_polyak_kernel)")
    ap.add_argument("--step", type=int, default=1, help="with --marker: which step (0-based) to summarise")
    a = ap.parse_args()
    lines = [l for l in open(a.csv, errors="replace") if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    name_c, val_c, unit_c, metric_c = col["Kernel Name"], col["Metric Value"], col["Metric Unit"], col["Metric Name"]
    launches = []
    for r in rows[1:]:
        if len(r) != len(hdr) or r[metric_c] != "gpu__time_duration.sum":
            continue
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}[r[unit_c]]
        name = re.sub(r"\(.*", "", r[name_c])
        name = name.replace("gac::", "")
        launches.append((name, float(r[val_c].replace(",", "")) * scale))
    if a.marker:
        ends = [i for i, (n, _) in enumerate(launches) if a.marker in n]
        lo = ends[a.step - 1] + 1 if a.step > 0 else 0
        launches = launches[lo:ends[a.step] + 1]
    launches = launches[a.skip:]
    if a.take:
        launches = launches[:a.take]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for n, t in launches:
        tot[n] += t
        cnt[n] += 1
    total = sum(tot.values())
    if a.title:
        print(f"# {a.title}")
    print(f"# {len(launches)} launches, {total / 1e3:.3f} ms summed (cold-cache, serialised under ncu: compare SHARES)")
    for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{100 * t / total:6.2f}% {t / 1e3:8.3f} ms n={cnt[n]:4d} avg {t / cnt[n]:8.1f} us  {n}")


if __name__ == "__main__":
    main()
