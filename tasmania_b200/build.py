# -*- coding: utf-8 -*-
"""Build ``libtasmania_b200.so`` in-tree with nvcc for sm_100a.

    python -m tasmania_b200.build [--force] [--verbose]

The shared library is self-contained (static cudart) and lands next to this file so that it
travels with the source tree; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libtasmania_b200.so")

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "--extended-lambda",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # parity with the reference's numpy arithmetic: no FMA contraction, IEEE div/sqrt
    "-fmad=false",
    "-prec-div=true",
    "-prec-sqrt=true",
    "-Xcompiler", "-fPIC",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "tasmania_b200.h"))
    return hs


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in sources() + headers() + [__file__])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    nvcc = nvcc_path()
    os.makedirs(BUILD, exist_ok=True)
    hdr_time = max(os.path.getmtime(p) for p in headers() + [__file__])

    def compile_one(src):
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time)):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", LIB, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ns = ap.parse_args()
    print(build(ns.force, ns.verbose))
    sys.exit(0)
