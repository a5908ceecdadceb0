"""Oracle restatement (oracle/tfhe_oracle.c) against golden vectors produced by the
reference's own code (tests/golden/make_golden.py).  CPU only."""
import ctypes
import os

import numpy as np
import pytest

from conftest import wrap32
from oracle.pyoracle import FFT_FOLDED, FFT_REF, GATES, Keys

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_integer_primitives_bit_exact(oracle):
    z = load("primitives.npz")
    for i, a in enumerate(z["rots"]):
        assert np.array_equal(oracle.mul_by_xai(int(a), z["poly"]), z["mul_by_xai"][i])
        assert np.array_equal(oracle.mul_by_xai(int(a), z["poly"], True), z["mul_by_xai_minus_one"][i])
    assert np.array_equal(oracle.decomp(z["poly"]), z["decomp"])
    got = [oracle.modswitch_from(int(p), 2048) for p in z["phases"]]
    assert np.array_equal(np.array(got, np.int32), z["modswitch_2048"])
    # closed form used by the CUDA kernel
    closed = ((z["phases"].astype(np.int64) & 0xFFFFFFFF) + (1 << 20)) % (1 << 32) >> 21
    assert np.array_equal(closed.astype(np.int32), z["modswitch_2048"])
    got = [oracle.modswitch_to(m, s) for m, s in ((1, 8), (-1, 8), (1, 4), (-1, 4))]
    assert np.array_equal(np.array(got, np.int32), z["modswitch_to"])


def _tiny_ctx(oracle, bk_rows, mode, ks=None):
    p = oracle.small_params(bk_rows.shape[0])
    ks = np.zeros((p.N * p.k, p.ks_t, 1 << p.ks_basebit, p.n + 1), np.int32) if ks is None else ks
    return oracle.ctx(Keys(p, np.zeros(p.n, np.int32), np.zeros(p.N, np.int32), bk_rows, ks), mode), p


def test_fourier_transforms(oracle):
    z = load("fft.npz")
    ctx, _ = _tiny_ctx(oracle, np.zeros((1, 4, 2, 1024), np.int32), FFT_REF)
    # same arithmetic as the reference build -> equal to rounding of the platform's libm
    assert np.allclose(ctx.ifft_int(z["small"]), z["ifft_int"], rtol=0, atol=1e-7)
    assert np.allclose(ctx.ifft_torus(z["poly"]), z["ifft_torus"], rtol=0, atol=1e-12)
    back = ctx.fft_torus(z["ifft_torus"])
    assert np.abs(wrap32(back.astype(np.int64) - z["fft_torus_of_ifft_torus"].astype(np.int64))).max() <= 1
    assert np.abs(wrap32(back.astype(np.int64) - z["poly"].astype(np.int64))).max() <= 1
    # folded representation: P(zeta^(4m+1)); m < 256 is the conjugate of reference value 2m,
    # m >= 256 is reference value 1023 - 2m  (include/tfhe_b200.h, tfhe_b200_load_bk_fourier)
    f, _ = _tiny_ctx(oracle, np.zeros((1, 4, 2, 1024), np.int32), FFT_FOLDED)
    mine = f.ifft_int(z["small"])
    m = np.arange(512)
    expect = np.where(m < 256, np.conj(z["ifft_int"][np.minimum(2 * m, 511)]),
                      z["ifft_int"][np.clip(1023 - 2 * m, 0, 511)])
    assert np.allclose(mine, expect, rtol=0, atol=1e-6)


@pytest.mark.parametrize("mode", [FFT_REF, FFT_FOLDED])
def test_external_product_and_blind_rotation(oracle, mode):
    z = load("blind_rotate.npz")
    ctx, p = _tiny_ctx(oracle, z["bk_rows"], mode)
    got = ctx.extern_mul(3, z["acc"])
    d = wrap32(got.astype(np.int64) - z["extern_mul_bk3"].astype(np.int64))
    assert np.abs(d).max() <= (0 if mode == FFT_REF else 2)
    exact = oracle.extern_mul_exact(z["bk_rows"][3], z["acc"], p)
    assert np.abs(wrap32(z["extern_mul_bk3"].astype(np.int64) - exact.astype(np.int64))).max() <= 1
    br = ctx.blind_rotate(z["acc0"], z["bara"])
    if mode == FFT_REF:
        assert np.array_equal(br, z["blind_rotate"])
    # phases under the TLWE key agree whatever the transform (6 iterations of key noise)
    key = z["tlwe_key"]

    def phase(acc):
        full = np.convolve(acc[0].astype(np.int64), key.astype(np.int64))
        r = full[:1024].copy()
        r[:1023] -= full[1024:]
        return wrap32(acc[1].astype(np.int64) - r)

    assert np.abs(wrap32(phase(br) - phase(z["blind_rotate"]))).max() / 2.0 ** 32 < 2.0 ** -8


def test_keyswitch_bit_exact(oracle):
    z = load("keyswitch.npz")
    p = oracle.small_params(500)
    ks = np.zeros((1024, 8, 4, 501), np.int32)
    ks[: z["ks_rows"].shape[0]] = z["ks_rows"]
    ctx = oracle.ctx(Keys(p, np.zeros(500, np.int32), np.zeros(1024, np.int32),
                          np.zeros((500, 4, 2, 1024), np.int32), ks), FFT_FOLDED)
    assert np.array_equal(ctx.keyswitch(z["u"]), z["out"])


def test_reference_gate_outputs_decrypt_to_truth_tables():
    """What the reference itself produced (seed {314,1592,657}): bits and noise margin."""
    z = load("gate_phases.npz")
    a = np.array([0, 0, 1, 1])
    b = np.array([0, 1, 0, 1])
    truth = {"NAND": 1 - (a & b), "OR": a | b, "AND": a & b, "XOR": a ^ b, "XNOR": 1 - (a ^ b),
             "NOR": 1 - (a | b), "ANDNY": (1 - a) & b, "ANDYN": a & (1 - b), "ORNY": (1 - a) | b,
             "ORYN": a | (1 - b)}
    for g in GATES:
        ph = z[g].astype(np.int64) / 2.0 ** 32
        assert np.array_equal((ph > 0).astype(int), truth[g]), g
        assert np.abs(np.abs(ph) - 0.125).max() < 2.0 ** -5
    ph = z["mux"].astype(np.int64) / 2.0 ** 32
    aa, bb, cc = np.meshgrid([0, 1], [0, 1], [0, 1], indexing="ij")
    assert np.array_equal((ph > 0).astype(int), np.where(aa == 1, bb, cc).reshape(-1))
