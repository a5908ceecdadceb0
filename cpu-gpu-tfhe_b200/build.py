#!/usr/bin/env python3
"""Builds cpu-gpu-tfhe_b200/libtfhe_b200.so (sm_100a only) with nvcc.

The library is built IN-TREE so that it travels with the repository snapshot to
the GPU box.  cudart is linked statically; there is no other dependency.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libtfhe_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["blind_rotate.cu", "keyswitch.cu", "keyswitch_mma.cu", "engine.cu", "client.cu", "microbench.cu", "compat.cu", "circuits.cu", "keyio.cu", "keygen.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall", "-I", os.path.join(HERE, "..", "include"),
]


def _deps():
    out = []
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                out.append(os.path.join(root, f))
    return out


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-6000:]))
    return r.stdout


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    gen = os.path.join(CSRC, "fft_consts.h")
    if not os.path.exists(gen):
        _run([sys.executable, os.path.join(CSRC, "gen_fft_consts.py")])
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    dep_time = max(os.path.getmtime(d) for d in _deps())
    jobs, objs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), dep_time):
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])
    if verbose and jobs:
        print("[build] compiling %d CUDA translation units for sm_100a" % len(jobs), flush=True)
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(_run, jobs))
    if force or jobs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
        if verbose:
            print("[build] linked", LIB, flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
