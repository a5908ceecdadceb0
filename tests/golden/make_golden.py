#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = the reference's own
host code compiled from /root/reference by oracle/build_ref.py).  Run in the build container:

    python tests/golden/make_golden.py

The reference ships no golden vectors; these pin the oracle restatement to
the reference's behaviour on machines where /root/reference (and oracle/_ref) is absent.
Full-size keys are ~80 MB, so the fixtures hold (i) primitives on explicit inputs and
(ii) results that depend on the key only through a few rows, together with those rows.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import GATES, Oracle, Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    o = Oracle()
    r = Ref().keygen((314, 1592, 657))  # the reference's own seed (main.cu:2724, cpu/main.cpp:21)
    keys = r.export_keys(o.params)
    rng = np.random.default_rng(2024)

    # ---- integer primitives --------------------------------------------------------------
    poly = rng.integers(-2 ** 31, 2 ** 31, 1024, dtype=np.int64).astype(np.int32)
    rots = np.array([1, 5, 16, 1023, 1024, 1500, 2047], np.int32)
    phases = rng.integers(-2 ** 31, 2 ** 31, 64, dtype=np.int64).astype(np.int32)
    np.savez_compressed(
        os.path.join(OUT, "primitives.npz"),
        poly=poly, rots=rots,
        mul_by_xai=np.stack([r.mul_by_xai(int(a), poly) for a in rots]),
        mul_by_xai_minus_one=np.stack([r.mul_by_xai(int(a), poly, True) for a in rots]),
        decomp=r.decomp(poly),
        phases=phases,
        modswitch_2048=np.array([r.modswitch_from(int(p), 2048) for p in phases], np.int32),
        modswitch_to=np.array([r.modswitch_to(m, s) for m, s in ((1, 8), (-1, 8), (1, 4), (-1, 4))], np.int32),
    )

    # ---- Fourier transforms (reference lagrangehalfc layout) -----------------------------------
    small = rng.integers(-512, 512, 1024).astype(np.int32)
    lag_int, lag_tor = r.ifft_int(small), r.ifft_torus(poly)
    np.savez_compressed(os.path.join(OUT, "fft.npz"), small=small, poly=poly, ifft_int=lag_int,
                        ifft_torus=lag_tor, fft_torus_of_ifft_torus=r.fft_torus(lag_tor))

    # ---- external product and a short blind rotation with the key rows they use ------------
    acc = rng.integers(-2 ** 31, 2 ** 31, (2, 1024), dtype=np.int64).astype(np.int32)
    n_iter = 6
    bara = np.array([3, 0, 2047, 1024, 517, 1], np.int32)
    acc0 = np.zeros((2, 1024), np.int32)
    acc0[1, :] = 1 << 29
    np.savez_compressed(
        os.path.join(OUT, "blind_rotate.npz"),
        bk_rows=keys.bk[:n_iter], tlwe_key=keys.tlwe_key, acc=acc, extern_mul_bk3=r.extern_mul(3, acc),
        bara=bara, acc0=acc0, blind_rotate=r.blind_rotate(acc0, bara),
    )

    # ---- key switch on a reduced input (only the first 8 mask words are non-zero) ------------
    u = np.zeros(1025, np.int32)
    u[:8] = rng.integers(-2 ** 31, 2 ** 31, 8, dtype=np.int64).astype(np.int32)
    # mask words equal to 0 still select digit rows through the rounding offset 2^15? No:
    # (0 + 2^15) >> (32 - 2(j+1)) & 3 == 0 for every j, so rows i >= 8 are never touched.
    u[1024] = 123456789
    np.savez_compressed(os.path.join(OUT, "keyswitch.npz"), ks_rows=keys.ks[:8], u=u, out=r.keyswitch(u))

    # ---- whole gates: inputs, decrypted outputs and output phases (keys too big to ship) ------
    bits = [(0, 0), (0, 1), (1, 0), (1, 1)]
    recs = {}
    for g in GATES:
        ph = []
        for a, b in bits:
            out = r.gate(g, r.encrypt(a), r.encrypt(b))
            ph.append(r.phase(out))
        recs[g] = np.array(ph, np.int32)
    mux = []
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                mux.append(r.phase(r.mux(r.encrypt(a), r.encrypt(b), r.encrypt(c))))
    np.savez_compressed(os.path.join(OUT, "gate_phases.npz"), mux=np.array(mux, np.int32), **recs)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
