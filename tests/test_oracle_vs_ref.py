"""Pins the oracle restatement to the reference's own host code (oracle/_ref, built from
/root/reference by oracle/build_ref.py): in FFT_REF mode every output is bit-identical.
Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

from oracle.pyoracle import FFT_REF, GATES, Oracle, Ref, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def pair():
    o = Oracle()
    r = Ref().keygen((314, 1592, 657))
    keys = r.export_keys(o.params)
    return o, r, keys, o.ctx(keys, FFT_REF)


def test_dimensions_and_noise_match_defaults(pair):
    o, r, _, _ = pair
    p = o.params
    assert (r.n, r.N, r.k, r.l, r.Bgbit, r.ks_t, r.ks_basebit) == (p.n, p.N, p.k, p.l, p.Bgbit, p.ks_t, p.ks_basebit)
    a = r.alphas()
    assert a[0] == p.alpha_lwe and a[1] == p.alpha_bk


def test_fourier_key_identical(pair):
    _, r, _, ctx = pair
    assert np.array_equal(ctx.bkfft(), r.bkfft())


def test_primitives_identical(pair):
    o, r, keys, ctx = pair
    rng = np.random.default_rng(5)
    poly = rng.integers(-2 ** 31, 2 ** 31, 1024, dtype=np.int64).astype(np.int32)
    assert np.array_equal(o.decomp(poly), r.decomp(poly))
    for a in (1, 7, 1023, 1024, 2047):
        assert np.array_equal(o.mul_by_xai(a, poly), r.mul_by_xai(a, poly))
        assert np.array_equal(o.mul_by_xai(a, poly, True), r.mul_by_xai(a, poly, True))
    for ph in rng.integers(-2 ** 31, 2 ** 31, 200, dtype=np.int64):
        assert o.modswitch_from(int(ph), 2048) == r.modswitch_from(int(ph), 2048)
    assert np.array_equal(ctx.ifft_int(poly >> 22), r.ifft_int(poly >> 22))
    assert np.array_equal(ctx.ifft_torus(poly), r.ifft_torus(poly))
    acc = rng.integers(-2 ** 31, 2 ** 31, (2, 1024), dtype=np.int64).astype(np.int32)
    assert np.array_equal(ctx.extern_mul(11, acc), r.extern_mul(11, acc))
    u = rng.integers(-2 ** 31, 2 ** 31, 1025, dtype=np.int64).astype(np.int32)
    assert np.array_equal(ctx.keyswitch(u), r.keyswitch(u))


def test_every_gate_and_mux_identical(pair):
    o, r, keys, ctx = pair
    ca, cb, cc = r.encrypt(1), r.encrypt(0), r.encrypt(1)
    for g in GATES:
        assert np.array_equal(ctx.gate(g, ca, cb), r.gate(g, ca, cb)), g
    assert np.array_equal(ctx.mux(ca, cb, cc), r.mux(ca, cb, cc))
    x = o.gate_prologue("AND", ca, cb)
    assert np.array_equal(ctx.bootstrap_woks(1 << 29, x), r.bootstrap_woks(1 << 29, x))


def test_oracle_keys_run_through_the_reference(pair):
    """Keys from the oracle's portable keygen are valid inputs for the reference code."""
    o, _, _, _ = pair
    keys = o.keygen(77)
    r2 = Ref().import_keys(keys)
    rng = o.rng(3)
    c = o.encrypt_bits(keys, rng, [1, 0])
    out = r2.gate("NAND", c[0], c[1])
    assert o.decrypt_bits(keys, out[None])[0] == 1
    assert np.array_equal(out, o.ctx(keys, FFT_REF).gate("NAND", c[0], c[1]))


def test_exact_blind_rotation_equals_reference_naive_path(pair):
    """The oracle's FFT-free blind rotation (oracle_blind_rotate_exact) against the reference's own
    non-FFT sequence — tLweMulByXaiMinusOne, tGswTLweDecompH, torusPolynomialMultNaive
    (multiplication.cu:72), tLweAddTo, i.e. tfhe_MuxRotate / tfhe_blindRotate of
    lwe-bootstrapping-functions.cu:34-79 with exact products — word for word; and the FFT path
    stays within 1 LSB per step of it (truncation of a value next to the exact integer)."""
    o, r, keys, ctx = pair
    rng = np.random.default_rng(9)
    acc = rng.integers(-2 ** 31, 2 ** 31, (2, 1024), dtype=np.int64).astype(np.int32)
    bara = np.array([5, 0, 2047, 1024, 1, 1023, 777, 1500], np.int32)
    exact = o.blind_rotate_exact(keys.bk, acc, bara)
    assert np.array_equal(exact, r.blind_rotate_naive(acc, bara))
    assert not np.array_equal(exact, acc)
    one = np.array([0, 0, 333], np.int32)      # a single step: FFT path = exact or exact -+ 1
    d = r.blind_rotate(acc, one).astype(np.int64) - o.blind_rotate_exact(keys.bk, acc, one).astype(np.int64)
    assert np.abs((d + 2 ** 31) % 2 ** 32 - 2 ** 31).max() <= 1


def test_exact_bootstrap_decrypts_and_matches_fft_bootstrap_bits(pair):
    """tfhe_bootstrap_woKS with exact products: same decrypted bit as the reference's FFT bootstrap,
    phase within bootstrap noise of it (the two differ in rounding only)."""
    o, r, keys, ctx = pair
    ca, cb = r.encrypt(1), r.encrypt(1)
    x = o.gate_prologue("NAND", ca, cb)
    u_exact = o.bootstrap_woks_exact(keys.bk, 1 << 29, x)[0]
    u_fft = r.bootstrap_woks(1 << 29, x)
    ke = np.concatenate([keys.tlwe_key.reshape(-1)])  # extracted key = TLWE key (lwe.cu:287-296)
    ph = lambda u: int(o.phases(ke, u[None])[0])
    assert ph(u_exact) < 0 and ph(u_fft) < 0          # NAND(1,1) = 0 -> -1/8
    assert abs(ph(u_exact) - ph(u_fft)) / 2.0 ** 32 < 2.0 ** -5
