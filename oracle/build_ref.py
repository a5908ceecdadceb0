#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — builds oracle/_ref/libtfhe_ref.so.

Compiles the reference's own host path from the sources WHERE THEY LIE under
/root/reference/gpuParallel (nothing is copied into this repo), links it with
the fp64 FFTW stand-in (oracle/fftw_shim) and the flat-array adapter
(oracle/ref_adapter.cpp).  The reference's own Makefile is not used (it targets
CUDA 10.1 / sm_30..sm_70 and has a malformed pattern rule).

Left out on purpose:
  main.cu, cloud.cu   -- CLI drivers with their own main()
  Cipher.cu           -- static initialiser fopen()s "cloud.key" at load time

Outputs go only to oracle/_ref/ (git-ignored, but shipped to the GPU box).
If /root/reference is absent (GPU box) the prebuilt .so is used as is.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TFHE_REFERENCE_DIR", "/root/reference/gpuParallel")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")
LIB = os.path.join(OUT, "libtfhe_ref.so")
SKIP = {"main.cu", "cloud.cu", "Cipher.cu"}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))
    return r.stdout


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=True):
    if not os.path.isdir(REF):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("reference sources not found at %s and no prebuilt %s" % (REF, LIB))
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(REF) if f.endswith(".cu") and f not in SKIP)
    common = [
        NVCC, "-std=c++14", "-w", "-O2", "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC,-fopenmp", "-I", os.path.join(HERE, "fftw_shim"),
        "-I", os.path.join(REF, "cuda_common", "inc"), "-I", REF,
    ]
    jobs = []
    for f in srcs:
        src = os.path.join(REF, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        if force or _newer(src, obj):
            jobs.append(common + ["-c", src, "-o", obj])
    adapter = os.path.join(HERE, "ref_adapter.cpp")
    adapter_o = os.path.join(OBJ, "ref_adapter.o")
    if force or _newer(adapter, adapter_o):
        jobs.append(common + ["-x", "cu", "-c", adapter, "-o", adapter_o])
    shim = os.path.join(HERE, "fftw_shim", "fftw_shim.c")
    shim_o = os.path.join(OBJ, "fftw_shim.o")
    if force or _newer(shim, shim_o):
        jobs.append(["gcc", "-O3", "-mavx2", "-mfma", "-fPIC", "-c", shim, "-o", shim_o])
    if verbose:
        print("[build_ref] compiling %d translation units" % len(jobs), flush=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        list(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, f[:-3] + ".o") for f in srcs] + [adapter_o, shim_o]
    if force or jobs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-lcufft", "-Xcompiler", "-fopenmp", "-lgomp"])
    if verbose:
        print("[build_ref] built", LIB, flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
