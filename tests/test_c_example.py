"""The drop-in boundary is a plain-C ABI: both headers compile as C99 with no CUDA / C++ / Python in
sight, and examples/cloud_add.c — the "cloud" side of the reference's client / cloud hand-off
(cpuParallel/cloud.cpp:138-161) written against include/tfhe_b200.h only — links against the library
with gcc.  On the GPU box the program is run end to end: the client (this package) writes cloud.key
and cloud.data, the C program computes sum and product on ciphertexts, the client decrypts answer.data."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "cpu-gpu-tfhe_b200")


def _build(tmp_path):
    exe = str(tmp_path / "cloud_add")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "cloud_add.c"), "-o", exe, "-L", LIBDIR, "-ltfhe_b200",
                    "-Wl,-rpath," + LIBDIR], check=True)
    return exe


@pytest.mark.parametrize("header", ["tfhe_b200.h", "tfhe_compat.h"])
def test_headers_are_plain_c99(header):
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(ROOT, "include", header)], check=True)


def test_c_example_links_and_fails_loudly_without_inputs(pkg, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "none.key"), "x", "y"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "cannot read" in r.stderr


@pytest.mark.gpu
def test_c_example_end_to_end(pkg, tmp_path):
    exe = _build(tmp_path)
    sk = pkg.keygen(77)
    nbits, a, b = 16, 12345, 54321
    bits = lambda v: ((np.array([v], np.int64)[:, None] >> np.arange(nbits)) & 1).astype(np.int32).reshape(-1)
    pkg.write_cloud_key(tmp_path / "cloud.key", sk)
    pkg.write_ciphertexts(tmp_path / "cloud.data", pkg.encrypt_bits(sk, bits(a), 1))
    pkg.write_ciphertexts(tmp_path / "cloud.data", pkg.encrypt_bits(sk, bits(b), 2), append=True)
    r = subprocess.run([exe, str(tmp_path / "cloud.key"), str(tmp_path / "cloud.data"), str(tmp_path / "answer.data"),
                        str(nbits)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "kernel launches" in r.stdout
    ans, _ = pkg.read_ciphertexts(tmp_path / "answer.data", sk.params.n)
    got = pkg.decrypt_bits(sk, ans).reshape(2, nbits).astype(np.int64)
    vals = (got << np.arange(nbits)).sum(-1)
    assert vals[0] == (a + b) % 2 ** nbits and vals[1] == (a * b) % 2 ** nbits
