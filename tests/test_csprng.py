"""The key owner's randomness (cpu-gpu-tfhe_b200/csrc/csprng.h): ChaCha20 keystream.  Known-answer
test of the block function against RFC 8439 section 2.3.2, OS-entropy keys differ from run to run,
seeded keys are reproducible (tests / benchmarks only), noise has the requested spread."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KAT = r'''
#include <cstdio>
#include "cpu-gpu-tfhe_b200/csrc/csprng.h"
int main() {
    tfhe_b200::ChaChaKey k;
    for (int i = 0; i < 8; i++) k.w[i] = (4u * i) | ((4u * i + 1) << 8) | ((4u * i + 2) << 16) | ((4u * i + 3) << 24);
    uint32_t out[16];
    // RFC 8439 2.3.2: counter = 1, nonce = 00:00:00:09 00:00:00:4a 00:00:00:00
    tfhe_b200::chacha20_block(k, 1ull | (0x09000000ull << 32), 0x4a000000ull, out);
    for (int i = 0; i < 16; i++) printf("%08x ", out[i]);
    printf("\n");
    return 0;
}
'''
RFC = ("e4e7f110 15593bd1 1fdd0f50 c47120a3 c7f4d1c7 0368c033 9aaa2204 4e6cd4c3 "
       "466482d2 09aa9f07 05d7c214 a2028bd9 d19c12b5 b94e16de e883d0cb 4e3c50a2")


def test_chacha20_block_matches_rfc8439(tmp_path):
    src = tmp_path / "kat.cpp"
    src.write_text(KAT)
    exe = tmp_path / "kat"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", ROOT, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.strip()
    assert out == RFC


def test_seeded_keys_reproducible_os_keys_fresh(pkg):
    p = pkg.default_params()
    p.n = 16  # small LWE dimension: key generation in milliseconds
    a, b = pkg.keygen(5, p), pkg.keygen(5, p)
    assert np.array_equal(a.lwe_key, b.lwe_key) and np.array_equal(a.bk, b.bk) and np.array_equal(a.ks, b.ks)
    c = pkg.keygen(6, p)
    assert not np.array_equal(a.bk, c.bk)
    x, y = pkg.keygen(0, p), pkg.keygen(0, p)      # seed 0: getrandom()
    assert not np.array_equal(x.tlwe_key, y.tlwe_key) and not np.array_equal(x.bk, y.bk)
    assert set(np.unique(x.lwe_key)) <= {0, 1} and 300 < int(x.tlwe_key.sum()) < 724


def test_encryption_noise_and_mask_statistics(pkg):
    sk = pkg.keygen(9)
    bits = np.zeros(4000, np.int32)
    c = pkg.encrypt_bits(sk, bits, 0)               # OS-entropy stream
    noise = pkg.phases(sk.lwe_key, c).astype(np.int64) / 2.0 ** 32 + 0.125
    assert abs(noise.std() / sk.alpha_lwe - 1.0) < 0.08 and abs(noise.mean()) < 4 * sk.alpha_lwe / 60
    mask = c[:, :-1].astype(np.int64)
    assert abs(mask.mean()) < 2 ** 31 * 0.005 and abs(mask.std() / (2 ** 32 / 12 ** 0.5) - 1.0) < 0.01
    assert not np.array_equal(c, pkg.encrypt_bits(sk, bits, 0))
    assert np.array_equal(pkg.encrypt_bits(sk, bits, 7), pkg.encrypt_bits(sk, bits, 7))
