"""Exact parity: the CUDA path against the reference's FFT-FREE bootstrap, word for word.

The reference has two blind rotations: the FFT one the gates use
(lwe-bootstrapping-functions-fft.cu:676-737) and the plain one
(tfhe_MuxRotate / tfhe_blindRotate, lwe-bootstrapping-functions.cu:34-79) whose polynomial
products are, by definition, torusPolynomialMultNaive (multiplication.cu:53-77): integer
arithmetic mod 2^32, no rounding.  The FFT path computes the same products in fp64 and then
TRUNCATES a value that lies within ~0.06 of the exact integer (fft_processor_fftw.cu:177), i.e.
it returns the exact product or the exact product -+ 1.  The CUDA path rounds to NEAREST instead
(br_core.cuh double_to_torus32), so it must return the exact product itself — in every one of
the 500 iterations, hence the complete bootstrap and, after the (integer) key switch, the
complete gate must equal the FFT-free path bit for bit.  tests/test_oracle_vs_ref.py pins the
oracle's exact path to the reference's own torusPolynomialMultNaive.

Measured margin (tests/test_host_emulation.py, the same phase functions on the CPU): the fp64
value is within 0.057 of an integer over 6*10^5 conversions (RMS 0.012), against the 0.5 that
would flip a rounding.
"""
import ctypes

import numpy as np
import pytest

from conftest import wrap32

pytestmark = pytest.mark.gpu

MU = 0x20000000
GATES = ["NAND", "OR", "AND", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]


def _rand_i32(rng, shape):
    return rng.integers(-2 ** 31, 2 ** 31, size=shape, dtype=np.int64).astype(np.int32)


def _rounds_to_nearest(engine):
    return engine.L.tfhe_b200_conversion_mode() == 0


def test_full_bootstrap_and_gate_bit_identical_to_exact_path(engine, oracle, keys, ctx_ref):
    """12 gates of all 10 types: GPU bootstrap without key switch (all 500 iterations) == the
    exact path word for word; GPU gate == key switch (integer) of the exact extracted sample."""
    assert _rounds_to_nearest(engine), "this test is for the default (round-to-nearest) build"
    rng = oracle.rng(77)
    r2 = np.random.default_rng(78)
    kinds = GATES + ["XOR", "NAND"]
    bits_a, bits_b = r2.integers(0, 2, len(kinds)).astype(np.int32), r2.integers(0, 2, len(kinds)).astype(np.int32)
    ca, cb = oracle.encrypt_bits(keys, rng, bits_a), oracle.encrypt_bits(keys, rng, bits_b)
    x = np.stack([oracle.gate_prologue(g, ca[i], cb[i]) for i, g in enumerate(kinds)])
    u_exact = oracle.bootstrap_woks_exact(keys.bk, MU, x)
    u_gpu = engine.bootstrap_woks(engine.to_device(x), MU).cpu().numpy()
    ndiff = int((u_gpu != u_exact).sum())
    assert ndiff == 0, "%d of %d extracted words differ from the exact path (max %d LSB)" % (
        ndiff, u_exact.size, np.abs(wrap32(u_gpu.astype(np.int64) - u_exact)).max())
    d_ca, d_cb = engine.to_device(ca), engine.to_device(cb)
    for i, g in enumerate(kinds):
        out = engine.gate(g, d_ca[i:i + 1].contiguous(), d_cb[i:i + 1].contiguous()).cpu().numpy()[0]
        assert np.array_equal(out, ctx_ref.keyswitch(u_exact[i])), g
    # and the exact path decrypts to the truth table (it is a correct bootstrap, not just "equal")
    table = {"NAND": 1 - (bits_a & bits_b), "OR": bits_a | bits_b, "AND": bits_a & bits_b, "XOR": bits_a ^ bits_b,
             "XNOR": 1 - (bits_a ^ bits_b), "NOR": 1 - (bits_a | bits_b), "ANDNY": (1 - bits_a) & bits_b,
             "ANDYN": bits_a & (1 - bits_b), "ORNY": (1 - bits_a) | bits_b, "ORYN": bits_a | (1 - bits_b)}
    outs = np.stack([ctx_ref.keyswitch(u_exact[i]) for i in range(len(kinds))])
    dec = oracle.decrypt_bits(keys, outs)
    assert all(dec[i] == table[g][i] for i, g in enumerate(kinds))


def test_mux_bit_identical_to_exact_path(engine, oracle, keys, ctx_ref):
    """bootsMUX (boot-gates.cu:407-448): two bootstraps without key switch, (1/8) + u1 + u2 in the
    extracted domain, one key switch — every word equal to the exact path."""
    assert _rounds_to_nearest(engine)
    rng = oracle.rng(79)
    a, b, c = (np.array(v, np.int32) for v in ([0, 1, 1, 0], [1, 0, 1, 0], [0, 1, 0, 1]))
    ea, eb, ec = (oracle.encrypt_bits(keys, rng, v) for v in (a, b, c))
    n = keys.params.n
    x1 = np.stack([oracle.gate_prologue("AND", ea[i], eb[i]) for i in range(4)])     # -1/8 + a + b
    x2 = np.stack([oracle.gate_prologue("ANDNY", ea[i], ec[i]) for i in range(4)])   # -1/8 - a + c
    u1, u2 = oracle.bootstrap_woks_exact(keys.bk, MU, x1), oracle.bootstrap_woks_exact(keys.bk, MU, x2)
    s = (u1.astype(np.int64) + u2.astype(np.int64))
    s[:, -1] += MU
    s = wrap32(s).astype(np.int32)
    expect = np.stack([ctx_ref.keyswitch(s[i]) for i in range(4)])
    got = engine.mux(engine.to_device(ea), engine.to_device(eb), engine.to_device(ec)).cpu().numpy()
    assert got.shape == (4, n + 1)
    assert np.array_equal(got, expect)
    assert np.array_equal(oracle.decrypt_bits(keys, got), np.where(a == 1, b, c))


def test_long_blind_rotation_random_accumulator_exact(engine, oracle, keys):
    """tfhe_blindRotate_FFT on RANDOM accumulators (full-range coefficients in both polynomials,
    the worst case for the fp64 products) over 64 iterations incl. skipped ones: == exact path."""
    assert _rounds_to_nearest(engine)
    rng = np.random.default_rng(80)
    count, n_iter = 6, 64
    acc = _rand_i32(rng, (count, 2, 1024))
    bara = rng.integers(0, 2048, size=(count, n_iter)).astype(np.int32)
    bara[rng.random((count, n_iter)) < 0.15] = 0
    got = engine.blind_rotate(engine.to_device(acc).clone(), engine.to_device(bara)).cpu().numpy()
    for r in range(count):
        assert np.array_equal(got[r], oracle.blind_rotate_exact(keys.bk, acc[r], bara[r])), r


def test_extern_mul_equals_exact(engine, oracle, keys):
    """tGswFFTExternMulToTLwe: == the exact product under round-to-nearest, +-1 LSB when the
    library was built with the reference's truncating conversion."""
    rng = np.random.default_rng(81)
    acc = _rand_i32(rng, (8, 2, 1024))
    bar = 0 if _rounds_to_nearest(engine) else 1
    for bk_index in (0, 2, 250, 499):
        got = engine.extern_mul(engine.to_device(acc).clone(), bk_index).cpu().numpy()
        for i in range(acc.shape[0]):
            exact = oracle.extern_mul_exact(keys.bk[bk_index], acc[i])
            assert np.abs(wrap32(got[i].astype(np.int64) - exact.astype(np.int64))).max() <= bar


def test_conversion_mode_is_reported(engine):
    engine.L.tfhe_b200_conversion_mode.restype = ctypes.c_int
    assert engine.L.tfhe_b200_conversion_mode() in (0, 1)


def test_truncating_build(pkg):
    """The second library of build(): libtfhe_b200_trunc.so = the same sources with
    -DTFHE_B200_TRUNCATE_LIKE_REFERENCE=1 (the reference's Torus32(int64_t(x)) conversion,
    fft_processor_fftw.cu:177).  Run in a subprocess (a process binds one library):
    +-1 LSB of the exact product, <= 1 LSB of the reference's FFT path, gates decrypt."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(os.path.dirname(pkg.lib_path()), "libtfhe_b200_trunc.so")
    assert os.path.exists(lib), "build() must produce the truncating variant"
    env = dict(os.environ, TFHE_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "trunc_variant_check.py")], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["ok"] and res["max_vs_exact"] <= 1 and res["differing_words"] > 0
