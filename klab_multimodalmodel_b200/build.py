"""Build libklab_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m klab_multimodalmodel_b200.build [--force] [--verbose]

Object files are cached under csrc/build/ keyed on a HASH of (source, every header, compiler flags) stored next to each
object, so a rebuild after touching one kernel only recompiles that translation unit and a stale cache under a fresh checkout
(mtimes say nothing there) is never reused.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libklab_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-diag-suppress", "177",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libklab_b200.so cannot be built (there is no CPU fallback)")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> bytes:
    hs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hs.append(os.path.join(os.path.dirname(HERE), "include", "klab_b200.h"))
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in hs:
        with open(path, "rb") as fh:
            h.update(path.encode() + b"\0" + fh.read())
    return h.digest()


def _stamp(src: str, hdr: bytes) -> str:
    with open(src, "rb") as fh:
        return hashlib.sha256(hdr + fh.read()).hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_digest()
    jobs = []
    objs = []
    stamps = {}
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        want = _stamp(src, hdr)
        have = open(obj + ".sha256").read().strip() if os.path.exists(obj + ".sha256") else ""
        stale = force or not os.path.exists(obj) or have != want
        if stale:
            cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
            stamps[obj] = want

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {cmd[-3]}")
                with open(cmd[-1] + ".sha256", "w") as fh:
                    fh.write(stamps[cmd[-1]])
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
