#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libgac_b200.so (cuobjdump -sass): the tcgen05 / TMA / bulk-copy evidence.
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
UTCHMMA = tcgen05.mma kind::f16, UTCQMMA/UTCIMMA other kinds, UTCHMMA.2CTA = cta_group::2, LDTM/STTM = tcgen05.ld/st,
UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dpo_replication_b200", "libgac_b200.so")
KEY = ("UTC", "LDTM", "STTM", "UBLKCP", "UTMA", "SYNCS", "HMMA", "IMMA", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "SHFL", "VOTE")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", n) for n in out]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            parts = op.split(".")
            base = parts[0]
            if base.startswith("UTC") and "2CTA" in parts:
                base += ".2CTA"
            cur[base] += 1
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(SO, ROOT)}: {len(kernels)} kernels; columns: total SASS instructions, then the opcodes of interest")
    for (mangled, c), name in sorted(zip(kernels.items(), names), key=lambda kv: -sum(kv[0][1].values())):
        sel = {k: v for k, v in c.items() if k.startswith(KEY)}
        body = " ".join(f"{k}={v}" for k, v in sorted(sel.items(), key=lambda kv: (-kv[1], kv[0])))
        print(f"{name.replace('gac::', ''):60s} {sum(c.values()):6d}  {body}")


if __name__ == "__main__":
    main()
