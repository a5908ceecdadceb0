"""Key and ciphertext FILE formats: the product's reader / writer against
the reference's own tfhe_io.cu compiled in oracle/_ref (cloud.key, secret.key, ciphertext
records), byte for byte, plus a committed golden fixture written by the reference."""
import filecmp
import os

import numpy as np
import pytest

from oracle.pyoracle import Ref, have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


def test_golden_ciphertext_file_written_by_the_reference(pkg):
    """tests/golden/ref_ciphertexts.bin + .npz were written by the reference (make_golden_files.py)."""
    g = np.load(os.path.join(GOLD, "ref_ciphertexts.npz"))
    path = os.path.join(GOLD, "ref_ciphertexts.bin")
    out, var = pkg.read_ciphertexts(path, 500)
    assert out.shape == g["samples"].shape and np.array_equal(out, g["samples"])
    assert np.array_equal(var, g["variances"])
    # header of a cloud key written by the reference
    sk, _ = None, None
    with open(os.path.join(GOLD, "ref_cloud_key_header.txt"), "rb") as f:
        head = f.read()
    assert head.startswith(b"-----BEGIN GATEBOOTSPARAMS-----\nks_basebit: 2\nks_t: 8\n-----END GATEBOOTSPARAMS-----\n")


def test_own_round_trip_and_header_text(pkg, tmp_path):
    sk = pkg.keygen(5)
    p = tmp_path / "cloud.key"
    pkg.write_cloud_key(p, sk)
    with open(os.path.join(GOLD, "ref_cloud_key_header.txt"), "rb") as f:
        head = f.read()
    with open(p, "rb") as f:
        assert f.read(len(head)) == head  # identical text sections for the default parameter set
    back, var = pkg.read_key_file(p)
    assert np.array_equal(back.bk, sk.bk) and np.array_equal(back.ks, sk.ks)
    assert [back.params.n, back.params.N, back.params.k, back.params.l, back.params.Bgbit, back.params.ks_t,
            back.params.ks_basebit] == [500, 1024, 1, 2, 10, 8, 2]
    s = tmp_path / "secret.key"
    pkg.write_secret_key(s, sk)
    back, _ = pkg.read_key_file(s, secret=True)
    assert np.array_equal(back.lwe_key, sk.lwe_key) and np.array_equal(back.tlwe_key, sk.tlwe_key)
    assert np.array_equal(back.bk, sk.bk)
    c = pkg.encrypt_bits(sk, np.array([1, 0, 1, 1, 0], np.int32), 3)
    pkg.write_ciphertexts(tmp_path / "cloud.data", c[:2])
    pkg.write_ciphertexts(tmp_path / "cloud.data", c[2:], append=True)
    got, var = pkg.read_ciphertexts(tmp_path / "cloud.data", 500)
    assert np.array_equal(got, c) and os.path.getsize(tmp_path / "cloud.data") == 5 * 2016


def test_malformed_files_are_rejected(pkg, tmp_path):
    bad = tmp_path / "bad.key"
    bad.write_bytes(b"-----BEGIN LWEPARAMS-----\nn: 500\n-----END LWEPARAMS-----\n")
    with pytest.raises(pkg.EngineError, match="GATEBOOTSPARAMS"):
        pkg.read_key_file(bad)
    with pytest.raises(pkg.EngineError, match="cannot open"):
        pkg.read_key_file(tmp_path / "missing.key")
    trunc = tmp_path / "trunc.data"
    trunc.write_bytes(b"\x2a\x00\x00\x00" + b"\x00" * 100)
    with pytest.raises(pkg.EngineError):
        pkg.read_ciphertexts(trunc, 500, count=1)


@needs_ref
def test_reference_reads_what_the_product_writes(pkg, tmp_path):
    sk = pkg.keygen(11)
    pkg.write_secret_key(tmp_path / "secret.key", sk)
    pkg.write_cloud_key(tmp_path / "cloud.key", sk)
    ref = Ref().read_secret_key(tmp_path / "secret.key")
    k = ref.export_keys(None)
    assert np.array_equal(k.lwe_key, sk.lwe_key) and np.array_equal(k.tlwe_key, sk.tlwe_key)
    assert np.array_equal(k.bk, sk.bk) and np.array_equal(k.ks, sk.ks)
    # and the reference decrypts the product's ciphertext file
    bits = np.array([1, 0, 0, 1, 1, 1, 0, 1], np.int32)
    c = pkg.encrypt_bits(sk, bits, 9)
    pkg.write_ciphertexts(tmp_path / "cloud.data", c)
    got, _ = ref.read_ciphertexts(tmp_path / "cloud.data", len(bits))
    assert np.array_equal(got, c)
    assert [int(ref.phase(s) > 0) for s in got] == list(bits)


@needs_ref
def test_product_reads_and_rewrites_reference_files_byte_for_byte(pkg, tmp_path):
    ref = Ref().keygen((7, 8, 9))
    ref.write_cloud_key(tmp_path / "cloud_ref.key")
    ref.write_secret_key(tmp_path / "secret_ref.key")
    k = ref.export_keys(None)
    ck, var = pkg.read_key_file(tmp_path / "cloud_ref.key")
    assert np.array_equal(ck.bk, k.bk) and np.array_equal(ck.ks, k.ks)
    sk, var2 = pkg.read_key_file(tmp_path / "secret_ref.key", secret=True)
    assert np.array_equal(sk.lwe_key, k.lwe_key) and np.array_equal(sk.tlwe_key, k.tlwe_key)
    # writing the same content back reproduces the reference's bytes exactly
    pkg.write_cloud_key(tmp_path / "cloud_mine.key", ck, variances=var)
    pkg.write_secret_key(tmp_path / "secret_mine.key", sk, variances=var2)
    assert filecmp.cmp(tmp_path / "cloud_ref.key", tmp_path / "cloud_mine.key", shallow=False)
    assert filecmp.cmp(tmp_path / "secret_ref.key", tmp_path / "secret_mine.key", shallow=False)
    # ciphertext records
    c = np.stack([ref.encrypt(b) for b in (1, 0, 1)])
    v = np.array([1e-9, 2e-9, 3e-9])
    ref.write_ciphertexts(tmp_path / "ref.data", c, v)
    pkg.write_ciphertexts(tmp_path / "mine.data", c, v)
    assert filecmp.cmp(tmp_path / "ref.data", tmp_path / "mine.data", shallow=False)
    got, gv = pkg.read_ciphertexts(tmp_path / "ref.data", 500)
    assert np.array_equal(got, c) and np.array_equal(gv, v)
