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
