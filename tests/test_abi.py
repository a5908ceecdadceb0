"""The C-ABI library loads on a machine without a GPU, exports every symbol declared in
include/*.h, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if not n.startswith("__")))


@pytest.mark.parametrize("header", [h for h in sorted(os.listdir(os.path.join(ROOT, "include"))) if h.endswith(".h")])
def test_every_declared_symbol_is_exported(pkg, header):
    L = ctypes.CDLL(pkg.lib_path())
    syms = declared_symbols(header)
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu tests")
    with pytest.raises(pkg.EngineError, match="no CUDA device|no CPU fallback"):
        pkg.Engine()


def test_bad_parameters_are_rejected(pkg):
    p = pkg.default_params()
    p.N = 2048
    with pytest.raises(pkg.EngineError, match="unsupported TGSW parameters|no CUDA device"):
        pkg.Engine(params=p)


def test_client_side_keygen_encrypt_decrypt(pkg, oracle):
    """Product-side keygen / encryption / decryption agree with the oracle's semantics."""
    from oracle.pyoracle import FFT_FOLDED, Keys

    sk = pkg.keygen(7)
    assert set(np.unique(sk.lwe_key)) <= {0, 1} and set(np.unique(sk.tlwe_key)) <= {0, 1}
    bits = np.array([0, 1, 1, 0, 1, 0, 0, 1], np.int32)
    c = pkg.encrypt_bits(sk, bits, 3)
    assert np.array_equal(pkg.decrypt_bits(sk, c), bits)
    assert np.array_equal(pkg.phases(sk.lwe_key, c), oracle.phases(sk.lwe_key, c))
    noise = pkg.phases(sk.lwe_key, c).astype(np.int64) / 2.0 ** 32 - np.where(bits == 1, 0.125, -0.125)
    assert np.abs(noise).max() < 8 * sk.alpha_lwe
    # the keys are valid TFHE keys: the oracle evaluates gates correctly with them
    keys = Keys(oracle.params, sk.lwe_key, sk.tlwe_key, sk.bk, sk.ks)
    ctx = oracle.ctx(keys, FFT_FOLDED)
    out = np.stack([ctx.gate("XOR", c[i], c[(i + 3) % 8]) for i in range(4)])
    assert np.array_equal(pkg.decrypt_bits(sk, out), bits[:4] ^ np.roll(bits, -3)[:4])


def test_host_container_helpers_roundtrip(pkg):
    """convertBitToNumber / convertNumberToBits / new_gate_bootstrapping_ciphertext_array
    (boot-gates.cu:513-556, tfhe_gate_bootstrapping.cu:93-108): host-only helpers of the batched API."""
    L = ctypes.CDLL(pkg.lib_path())
    vp = ctypes.c_void_p

    class LweParams(ctypes.Structure):
        _fields_ = [("n", ctypes.c_int), ("alpha_min", ctypes.c_double), ("alpha_max", ctypes.c_double)]

    class ParamSet(ctypes.Structure):
        _fields_ = [("ks_t", ctypes.c_int), ("ks_basebit", ctypes.c_int), ("in_out_params", vp), ("tgsw_params", vp)]

    class CloudKeySet(ctypes.Structure):
        _fields_ = [("params", vp), ("bk", vp), ("bkFFT", vp)]

    class LweSample(ctypes.Structure):
        _fields_ = [("a", ctypes.POINTER(ctypes.c_int32)), ("b", ctypes.c_int32), ("current_variance", ctypes.c_double)]

    class LweSample16(ctypes.Structure):
        _fields_ = [("a", ctypes.POINTER(ctypes.c_int32)), ("b", ctypes.POINTER(ctypes.c_int32)),
                    ("current_variance", ctypes.POINTER(ctypes.c_double))]

    n, bits = 500, 5
    lp = LweParams(n, 1e-5, 1e-2)
    ps = ParamSet(8, 2, ctypes.addressof(lp), None)
    ck = CloudKeySet(ctypes.addressof(ps), None, None)
    L.new_gate_bootstrapping_ciphertext_array.restype = ctypes.POINTER(LweSample)
    L.new_gate_bootstrapping_ciphertext_array.argtypes = [ctypes.c_int, vp]
    L.convertBitToNumber.restype = ctypes.POINTER(LweSample16)
    L.convertBitToNumber.argtypes = [ctypes.POINTER(LweSample), ctypes.c_int, vp]
    L.convertNumberToBits.restype = ctypes.POINTER(LweSample)
    L.convertNumberToBits.argtypes = [ctypes.POINTER(LweSample16), ctypes.c_int, vp]
    arr = L.new_gate_bootstrapping_ciphertext_array(bits, ctypes.addressof(ps))
    rng = np.random.default_rng(0)
    ref = rng.integers(-2 ** 31, 2 ** 31, (bits, n + 1), dtype=np.int64).astype(np.int32)
    for i in range(bits):
        for j in range(n):
            arr[i].a[j] = int(ref[i, j])
        arr[i].b = int(ref[i, n])
        arr[i].current_variance = 0.5 * i
    num = L.convertBitToNumber(arr, bits, ctypes.addressof(ck))
    flat = np.ctypeslib.as_array(num.contents.a, shape=(bits * n,)).reshape(bits, n)
    assert np.array_equal(flat, ref[:, :n])
    assert [num.contents.b[i] for i in range(bits)] == [int(x) for x in ref[:, n]]
    back = L.convertNumberToBits(num, bits, ctypes.addressof(ck))
    for i in range(bits):
        assert back[i].b == int(ref[i, n]) and back[i].current_variance == 0.5 * i
        assert [back[i].a[j] for j in (0, 1, n - 1)] == [int(ref[i, j]) for j in (0, 1, n - 1)]
    L.freeLweSample_16.argtypes = [ctypes.POINTER(LweSample16)]
    L.freeLweSample_16(num)
    L.delete_gate_bootstrapping_ciphertext_array.argtypes = [ctypes.c_int, ctypes.POINTER(LweSample)]
    L.delete_gate_bootstrapping_ciphertext_array(bits, arr)
    L.delete_gate_bootstrapping_ciphertext_array(bits, back)
