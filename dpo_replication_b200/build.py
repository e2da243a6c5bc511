"""Build libgac_b200.so in-tree with nvcc for sm_100a (run: python -m dpo_replication_b200.build)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "gac_b200.cu")
OUT = os.path.join(HERE, "libgac_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
         "-I", os.path.join(ROOT, "include")]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(ROOT, "include", "gac_b200.h")]


def up_to_date() -> bool:
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile when the library is missing or older than a source.  Safe under concurrent callers (one process per GPU all
    call this first): an exclusive file lock serialises them, the late ones find the library up to date, and the compiler
    writes to a temporary name that is renamed into place, so no process ever dlopens a half-written file."""
    if not force and up_to_date():
        return OUT
    import fcntl
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and up_to_date():
                return OUT
            tmp = f"{OUT}.tmp.{os.getpid()}"
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", tmp]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libgac_b200.so")
            if verbose:
                sys.stderr.write(r.stderr)
            os.replace(tmp, OUT)
            return OUT
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
