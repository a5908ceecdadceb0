"""TEST INFRASTRUCTURE — ctypes bindings for the CPU oracle.

Two libraries are wrapped:

* ``Oracle``  -> oracle/liboracle.so   (plain-C restatement, oracle/tfhe_oracle.c)
* ``Ref``     -> oracle/_ref/libtfhe_ref.so (the reference's own host code built
  by oracle/build_ref.py from /root/reference, plus oracle/ref_adapter.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing here is on the product path.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtfhe_ref.so")

GATES = ["NAND", "OR", "AND", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]
GATE_ID = {g: i for i, g in enumerate(GATES)}
FFT_REF, FFT_FOLDED = 0, 1

_vp = ctypes.c_void_p


def _p(a):
    return a.ctypes.data_as(_vp)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build_oracle(force=False):
    """Compile oracle/liboracle.so with the committed Makefile."""
    if force or not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(ORACLE_SO)
        for f in ("tfhe_oracle.c", "tfhe_oracle.h", "fftw_shim/fftw_shim.c", "Makefile")
    ):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return ORACLE_SO


def have_ref():
    return os.path.exists(REF_SO)


class OracleParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("n", "N", "k", "l", "Bgbit", "ks_t", "ks_basebit")] + [
        ("alpha_lwe", ctypes.c_double), ("alpha_bk", ctypes.c_double)]


class OracleRng(ctypes.Structure):
    _fields_ = [("s", ctypes.c_uint64 * 4), ("has_spare", ctypes.c_int), ("spare", ctypes.c_double)]


class Keys:
    """Flat key material (see tfhe_oracle.h for the layouts)."""

    def __init__(self, params, lwe_key, tlwe_key, bk, ks):
        self.params, self.lwe_key, self.tlwe_key, self.bk, self.ks = params, lwe_key, tlwe_key, bk, ks


class Oracle:
    def __init__(self):
        build_oracle()
        L = self.L = ctypes.CDLL(ORACLE_SO)
        L.oracle_ctx_new.restype = _vp
        L.oracle_ctx_bkfft.restype = ctypes.POINTER(ctypes.c_double)
        L.oracle_bk_words.restype = ctypes.c_size_t
        L.oracle_ks_words.restype = ctypes.c_size_t
        L.oracle_lwe_phase.restype = ctypes.c_int32
        L.oracle_modswitch_to.restype = ctypes.c_int32
        L.oracle_dtot32.restype = ctypes.c_int32
        L.oracle_dtot32.argtypes = [ctypes.c_double]
        L.oracle_rng_torus.restype = ctypes.c_int32
        L.oracle_rng_gauss.restype = ctypes.c_double
        L.oracle_rng_gauss.argtypes = [_vp, ctypes.c_double]
        L.oracle_lwe_encrypt.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int32, ctypes.c_double, _vp]
        L.oracle_keygen.argtypes = [_vp, ctypes.c_uint64, _vp, _vp, _vp, _vp]
        L.oracle_rng_seed.argtypes = [_vp, ctypes.c_uint64]
        self.params = OracleParams()
        L.oracle_default_params(ctypes.byref(self.params))

    # -- parameters / keys ------------------------------------------------
    def small_params(self, n):
        """Default parameter set with a shorter LWE dimension (test fixtures)."""
        p = OracleParams()
        self.L.oracle_default_params(ctypes.byref(p))
        p.n = n
        return p

    def keygen(self, seed, params=None):
        p = params or self.params
        lwe = np.zeros(p.n, np.int32)
        tlwe = np.zeros(p.k * p.N, np.int32)
        bk = np.zeros(self.L.oracle_bk_words(ctypes.byref(p)), np.int32)
        ks = np.zeros(self.L.oracle_ks_words(ctypes.byref(p)), np.int32)
        self.L.oracle_keygen(ctypes.byref(p), seed, _p(lwe), _p(tlwe), _p(bk), _p(ks))
        kpl = (p.k + 1) * p.l
        return Keys(p, lwe, tlwe, bk.reshape(p.n, kpl, p.k + 1, p.N),
                    ks.reshape(p.N * p.k, p.ks_t, 1 << p.ks_basebit, p.n + 1))

    def rng(self, seed):
        r = OracleRng()
        self.L.oracle_rng_seed(ctypes.byref(r), seed)
        return r

    def encrypt_bits(self, keys, rng, bits):
        p = keys.params
        bits = np.asarray(bits).reshape(-1)
        out = np.zeros((bits.size, p.n + 1), np.int32)
        for i, b in enumerate(bits):
            self.L.oracle_encrypt_bit(ctypes.byref(p), ctypes.byref(rng), _p(keys.lwe_key), int(b), _p(out[i]))
        return out

    def phase(self, key, sample):
        sample = _i32(sample)
        return int(self.L.oracle_lwe_phase(_p(sample), _p(key), key.size))

    def phases(self, key, samples):
        samples = _i32(samples)
        a = samples[..., :-1].astype(np.uint32)
        b = samples[..., -1].astype(np.uint32)
        with np.errstate(over="ignore"):
            dot = (a * key.astype(np.uint32)).sum(axis=-1, dtype=np.uint32)
            return (b - dot).astype(np.int32)

    def decrypt_bits(self, keys, samples):
        return (self.phases(keys.lwe_key, samples) > 0).astype(np.int32)

    # -- numeric / polynomial primitives ----------------------------------
    def modswitch_from(self, phase, msize):
        return int(self.L.oracle_modswitch_from(ctypes.c_int32(int(phase)), msize))

    def modswitch_to(self, mu, msize):
        return int(self.L.oracle_modswitch_to(mu, msize))

    def mul_by_xai(self, a, poly, minus_one=False):
        poly = _i32(poly)
        out = np.zeros_like(poly)
        f = self.L.oracle_mul_by_xai_minus_one if minus_one else self.L.oracle_mul_by_xai
        f(int(a), poly.size, _p(poly), _p(out))
        return out

    def decomp(self, poly, params=None):
        p = params or self.params
        poly = _i32(poly)
        out = np.zeros((p.l, p.N), np.int32)
        self.L.oracle_decomp(ctypes.byref(p), _p(poly), _p(out))
        return out

    def gate_prologue(self, gate, ca, cb, params=None):
        p = params or self.params
        ca, cb = _i32(ca), _i32(cb)
        x = np.zeros(p.n + 1, np.int32)
        self.L.oracle_gate_prologue(ctypes.byref(p), GATE_ID[gate], _p(ca), _p(cb), _p(x))
        return x

    # -- context -----------------------------------------------------------
    def ctx(self, keys, fft_mode=FFT_REF):
        return OracleCtx(self, keys, fft_mode)

    def extern_mul_exact(self, bk_i, accum, params=None):
        p = params or self.params
        acc = _i32(accum).copy()
        bk_i = _i32(bk_i)
        self.L.oracle_extern_mul_exact(ctypes.byref(p), _p(bk_i), _p(acc))
        return acc


    def blind_rotate_exact(self, bk, accum, bara, n_iter=None, params=None):
        """tfhe_blindRotate with exact products (no FFT): the integer answer."""
        p = params or self.params
        acc, bara, bk = _i32(accum).copy(), _i32(bara), _i32(bk)
        self.L.oracle_blind_rotate_exact(ctypes.byref(p), _p(bk), _p(acc), _p(bara),
                                         int(bara.size if n_iter is None else n_iter))
        return acc

    def bootstrap_woks_exact(self, bk, mu, x, threads=0, params=None):
        """tfhe_bootstrap_woKS with exact products on a batch x[count][n+1] -> u[count][N+1]."""
        p = params or self.params
        x, bk = np.atleast_2d(_i32(x)), _i32(bk)
        u = np.zeros((x.shape[0], p.N * p.k + 1), np.int32)
        self.L.oracle_bootstrap_woks_exact_batch(ctypes.byref(p), _p(bk), ctypes.c_int32(int(mu)), _p(x), _p(u),
                                                 int(x.shape[0]), int(threads))
        return u


class OracleCtx:
    def __init__(self, oracle, keys, fft_mode):
        self.o, self.L, self.keys, self.p = oracle, oracle.L, keys, keys.params
        self._bk = _i32(keys.bk)
        self._ks = _i32(keys.ks)
        self.h = _vp(self.L.oracle_ctx_new(ctypes.byref(self.p), _p(self._bk), _p(self._ks), fft_mode))

    def __del__(self):
        try:
            self.L.oracle_ctx_free(self.h)
        except Exception:
            pass

    def bkfft(self):
        p = self.p
        kpl = (p.k + 1) * p.l
        n = p.n * kpl * (p.k + 1) * (p.N // 2) * 2
        arr = np.ctypeslib.as_array(self.L.oracle_ctx_bkfft(self.h), shape=(n,))
        return arr.view(np.complex128).reshape(p.n, kpl, p.k + 1, p.N // 2)

    def ifft_int(self, poly):
        poly = _i32(poly)
        out = np.zeros(self.p.N // 2, np.complex128)
        self.L.oracle_ifft_int(self.h, _p(poly), _p(out))
        return out

    def ifft_torus(self, poly):
        poly = _i32(poly)
        out = np.zeros(self.p.N // 2, np.complex128)
        self.L.oracle_ifft_torus(self.h, _p(poly), _p(out))
        return out

    def fft_torus(self, lag):
        lag = np.ascontiguousarray(lag, dtype=np.complex128)
        out = np.zeros(self.p.N, np.int32)
        self.L.oracle_fft_torus(self.h, _p(lag), _p(out))
        return out

    def extern_mul(self, bk_index, accum):
        acc = _i32(accum).copy()
        self.L.oracle_extern_mul(self.h, int(bk_index), _p(acc))
        return acc

    def blind_rotate(self, accum, bara, n_iter=None):
        acc = _i32(accum).copy()
        bara = _i32(bara)
        self.L.oracle_blind_rotate(self.h, _p(acc), _p(bara), int(bara.size if n_iter is None else n_iter))
        return acc

    def blind_rotate_and_extract(self, testvect, barb, bara):
        tv, bara = _i32(testvect), _i32(bara)
        u = np.zeros(self.p.N * self.p.k + 1, np.int32)
        self.L.oracle_blind_rotate_and_extract(self.h, _p(tv), int(barb), _p(bara), int(bara.size), _p(u))
        return u

    def bootstrap_woks(self, mu, x):
        x = _i32(x)
        u = np.zeros(self.p.N * self.p.k + 1, np.int32)
        self.L.oracle_bootstrap_woks(self.h, ctypes.c_int32(int(mu)), _p(x), _p(u))
        return u

    def keyswitch(self, u):
        u = _i32(u)
        out = np.zeros(self.p.n + 1, np.int32)
        self.L.oracle_keyswitch(self.h, _p(u), _p(out))
        return out

    def gate(self, gate, ca, cb):
        ca, cb = _i32(ca), _i32(cb)
        out = np.zeros(self.p.n + 1, np.int32)
        self.L.oracle_gate(self.h, GATE_ID[gate], _p(ca), _p(cb), _p(out))
        return out

    def gate_batch(self, gate, ca, cb, threads=0):
        ca, cb = _i32(ca), _i32(cb)
        out = np.zeros_like(ca)
        self.L.oracle_gate_batch(self.h, GATE_ID[gate], _p(ca), _p(cb), _p(out), int(ca.shape[0]), int(threads))
        return out

    def mux(self, a, b, c):
        a, b, c = _i32(a), _i32(b), _i32(c)
        out = np.zeros(self.p.n + 1, np.int32)
        self.L.oracle_mux(self.h, _p(a), _p(b), _p(c), _p(out))
        return out

    def add(self, a, b):
        a, b = _i32(a), _i32(b)
        out = np.zeros_like(a)
        self.L.oracle_add(self.h, _p(a), _p(b), int(a.shape[0]), _p(out))
        return out

    def mul(self, a, b):
        a, b = _i32(a), _i32(b)
        out = np.zeros_like(a)
        self.L.oracle_mul(self.h, _p(a), _p(b), int(a.shape[0]), _p(out))
        return out


    def cipher_mul(self, a, b, threads=1):
        """Cipher operator* of the CPU reference (2*nbits-bit product, OpenMP reduction over `threads`)."""
        a, b = _i32(a), _i32(b)
        nbits = a.shape[0]
        out = np.zeros((2 * nbits, a.shape[1]), np.int32)
        self.L.oracle_cipher_mul(self.h, _p(a), _p(b), int(nbits), int(threads), _p(out))
        return out

    def matmul_units(self, a, b, cacc, threads=1):
        """units x (temp = a*b; c += temp), the inner-loop body of cpuParallel/cloud.cpp:390-408."""
        a, b = _i32(a), _i32(b)
        cacc = _i32(cacc).copy()
        units, nbits = a.shape[0], a.shape[1]
        self.L.oracle_matmul_units(self.h, _p(a), _p(b), _p(cacc), int(nbits), int(units), int(threads))
        return cacc


class Ref:
    """The reference's own host path (default parameter set only)."""

    def __init__(self):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libtfhe_ref.so missing: run oracle/build_ref.py")
        L = self.L = ctypes.CDLL(REF_SO)
        L.ref_keygen.restype = _vp
        L.ref_import.restype = _vp
        L.ref_phase.restype = ctypes.c_int32
        L.ref_modswitch_to.restype = ctypes.c_int32
        L.ref_time_nand.restype = ctypes.c_double
        d = (ctypes.c_int * 7)()
        L.ref_dims(d)
        self.n, self.N, self.k, self.l, self.Bgbit, self.ks_t, self.ks_basebit = list(d)
        self.kpl = (self.k + 1) * self.l
        self.h = None

    def alphas(self):
        a = (ctypes.c_double * 3)()
        self.L.ref_alphas(a)
        return list(a)

    def keygen(self, seed=(314, 1592, 657)):
        s = (ctypes.c_uint32 * len(seed))(*seed)
        self.h = _vp(self.L.ref_keygen(s, len(seed)))
        return self

    def reseed(self, seed):
        s = (ctypes.c_uint32 * len(seed))(*seed)
        self.L.ref_reseed(s, len(seed))

    def import_keys(self, keys):
        self._imp = (_i32(keys.lwe_key), _i32(keys.tlwe_key), _i32(keys.bk), _i32(keys.ks))
        self.h = _vp(self.L.ref_import(*[_p(a) for a in self._imp]))
        return self

    def export_keys(self, params):
        lwe = np.zeros(self.n, np.int32)
        tlwe = np.zeros(self.k * self.N, np.int32)
        bk = np.zeros((self.n, self.kpl, self.k + 1, self.N), np.int32)
        ks = np.zeros((self.N * self.k, self.ks_t, 1 << self.ks_basebit, self.n + 1), np.int32)
        self.L.ref_export_lwe_key(self.h, _p(lwe))
        self.L.ref_export_tlwe_key(self.h, _p(tlwe))
        self.L.ref_export_bk(self.h, _p(bk))
        self.L.ref_export_ks(self.h, _p(ks))
        return Keys(params, lwe, tlwe, bk, ks)

    def bkfft(self):
        out = np.zeros((self.n, self.kpl, self.k + 1, self.N // 2), np.complex128)
        self.L.ref_export_bkfft(self.h, _p(out))
        return out

    # ---- the reference's own file formats (tfhe_io.cu) ----
    def write_cloud_key(self, path):
        assert self.L.ref_write_cloud_key(self.h, str(path).encode()) == 0

    def write_secret_key(self, path):
        assert self.L.ref_write_secret_key(self.h, str(path).encode()) == 0

    def read_cloud_key(self, path):
        self.L.ref_read_cloud_key.restype = _vp
        self.h = _vp(self.L.ref_read_cloud_key(str(path).encode()))
        assert self.h
        return self

    def read_secret_key(self, path):
        self.L.ref_read_secret_key.restype = _vp
        self.h = _vp(self.L.ref_read_secret_key(str(path).encode()))
        assert self.h
        return self

    def write_ciphertexts(self, path, samples, variances=None):
        s = _i32(samples).reshape(-1, self.n + 1)
        v = None if variances is None else np.ascontiguousarray(variances, np.float64)
        assert self.L.ref_write_ciphertexts(self.h, str(path).encode(), _p(s), None if v is None else _p(v),
                                            s.shape[0]) == 0

    def read_ciphertexts(self, path, count):
        out = np.zeros((count, self.n + 1), np.int32)
        var = np.zeros(count, np.float64)
        assert self.L.ref_read_ciphertexts(self.h, str(path).encode(), _p(out), _p(var), count) == 0
        return out, var

    def encrypt(self, bit):
        out = np.zeros(self.n + 1, np.int32)
        self.L.ref_encrypt(self.h, int(bit), _p(out))
        return out

    def phase(self, sample):
        sample = _i32(sample)
        return int(self.L.ref_phase(self.h, _p(sample)))

    def gate(self, gate, ca, cb):
        ca, cb = _i32(ca), _i32(cb)
        out = np.zeros(self.n + 1, np.int32)
        self.L.ref_gate(self.h, GATE_ID[gate], _p(ca), _p(cb), _p(out))
        return out

    def mux(self, a, b, c):
        a, b, c = _i32(a), _i32(b), _i32(c)
        out = np.zeros(self.n + 1, np.int32)
        self.L.ref_mux(self.h, _p(a), _p(b), _p(c), _p(out))
        return out

    def bootstrap_woks(self, mu, x):
        x = _i32(x)
        u = np.zeros(self.N * self.k + 1, np.int32)
        self.L.ref_bootstrap_woks(self.h, ctypes.c_int32(int(mu)), _p(x), _p(u))
        return u

    def keyswitch(self, u):
        u = _i32(u)
        out = np.zeros(self.n + 1, np.int32)
        self.L.ref_keyswitch(self.h, _p(u), _p(out))
        return out

    def extern_mul(self, bk_index, accum):
        acc = _i32(accum).copy()
        self.L.ref_extern_mul(self.h, int(bk_index), _p(acc))
        return acc

    def blind_rotate(self, accum, bara, n_iter=None):
        acc = _i32(accum).copy()
        bara = _i32(bara)
        self.L.ref_blind_rotate(self.h, _p(acc), _p(bara), int(bara.size if n_iter is None else n_iter))
        return acc

    def blind_rotate_naive(self, accum, bara, n_iter=None):
        """The reference's non-FFT blind rotation with torusPolynomialMultNaive products."""
        acc = _i32(accum).copy()
        bara = _i32(bara)
        self.L.ref_blind_rotate_naive(self.h, _p(acc), _p(bara), int(bara.size if n_iter is None else n_iter))
        return acc

    def decomp(self, poly):
        poly = _i32(poly)
        out = np.zeros((self.l, self.N), np.int32)
        self.L.ref_decomp(self.h, _p(poly), _p(out))
        return out

    def mul_by_xai(self, a, poly, minus_one=False):
        poly = _i32(poly)
        out = np.zeros_like(poly)
        self.L.ref_mul_by_xai(int(a), int(bool(minus_one)), poly.size, _p(poly), _p(out))
        return out

    def modswitch_from(self, phase, msize):
        return int(self.L.ref_modswitch_from(ctypes.c_int32(int(phase)), msize))

    def modswitch_to(self, mu, msize):
        return int(self.L.ref_modswitch_to(mu, msize))

    def ifft_int(self, poly):
        poly = _i32(poly)
        out = np.zeros(self.N // 2, np.complex128)
        self.L.ref_ifft_int(_p(poly), _p(out))
        return out

    def ifft_torus(self, poly):
        poly = _i32(poly)
        out = np.zeros(self.N // 2, np.complex128)
        self.L.ref_ifft_torus(_p(poly), _p(out))
        return out

    def fft_torus(self, lag):
        lag = np.ascontiguousarray(lag, dtype=np.complex128)
        out = np.zeros(self.N, np.int32)
        self.L.ref_fft_torus(_p(lag), _p(out))
        return out

    def time_nand(self, ca, cb, count):
        ca, cb = _i32(ca), _i32(cb)
        last = np.zeros(self.n + 1, np.int32)
        return float(self.L.ref_time_nand(self.h, _p(ca), _p(cb), int(ca.shape[0]), int(count), _p(last))), last

    def close(self):
        if self.h:
            self.L.ref_free(self.h)
            self.h = None
