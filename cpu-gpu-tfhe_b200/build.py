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
# The same library with the reference's truncating fp64 -> Torus32 conversion
# (-DTFHE_B200_TRUNCATE_LIKE_REFERENCE=1, br_core.cuh double_to_torus32); only blind_rotate.cu differs.
LIB_TRUNC = os.path.join(HERE, "libtfhe_b200_trunc.so")
INCLUDE = os.path.join(HERE, "..", "include")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["blind_rotate.cu", "keyswitch.cu", "keyswitch_mma.cu", "engine.cu", "client.cu", "microbench.cu", "compat.cu", "circuits.cu", "keyio.cu", "keygen.cu", "multi.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas", "-I", os.path.join(HERE, "..", "include"),
]


def exported_symbols():
    """The C ABI = every function declared in include/*.h (nothing else leaves the library)."""
    import re

    names = set()
    for h in sorted(os.listdir(INCLUDE)):
        if not h.endswith(".h"):
            continue
        text = open(os.path.join(INCLUDE, h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        names |= set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text))
    return sorted(n for n in names if not n.startswith("__"))


def _version_script():
    path = os.path.join(OBJ, "exports.map")
    text = "{\n  global:\n" + "".join("    %s;\n" % n for n in exported_symbols()) + \
           "    tfhe_b200_debug_*;\n  local: *;\n};\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


def _link(lib, objs):
    _run([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                 "-Xlinker", "--version-script=" + _version_script()])


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
    # truncating-conversion variant: one more object
    br_src = os.path.join(CSRC, "blind_rotate.cu")
    br_trunc = os.path.join(OBJ, "blind_rotate_trunc.o")
    trunc_stale = force or not os.path.exists(br_trunc) or os.path.getmtime(br_trunc) < max(os.path.getmtime(br_src), dep_time)
    if trunc_stale:
        jobs.append([NVCC] + FLAGS + ["-DTFHE_B200_TRUNCATE_LIKE_REFERENCE=1", "-c", br_src, "-o", br_trunc])
    if verbose and jobs:
        print("[build] compiling %d CUDA translation units for sm_100a" % len(jobs), flush=True)
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(_run, jobs))
    hdr_time = max(os.path.getmtime(os.path.join(INCLUDE, h)) for h in os.listdir(INCLUDE))
    if force or jobs or not os.path.exists(LIB) or os.path.getmtime(LIB) < hdr_time:
        _link(LIB, objs)
        if verbose:
            print("[build] linked", LIB, flush=True)
    if force or jobs or not os.path.exists(LIB_TRUNC) or os.path.getmtime(LIB_TRUNC) < hdr_time:
        _link(LIB_TRUNC, [br_trunc if o.endswith("blind_rotate.o") else o for o in objs])
        if verbose:
            print("[build] linked", LIB_TRUNC, flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
