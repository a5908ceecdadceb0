"""Run by tests/test_gpu_exact.py::test_truncating_build in a subprocess with
TFHE_B200_LIB=cpu-gpu-tfhe_b200/libtfhe_b200_trunc.so: the library built with
-DTFHE_B200_TRUNCATE_LIKE_REFERENCE=1, whose fp64 -> Torus32 conversion is the reference's
Torus32(int64_t(x)) (fft_processor_fftw.cu:177).  Checks, on the GPU:
  * the build reports conversion mode 1;
  * one external product / MuxRotate step is within +-1 LSB of the EXACT integer product and
    within 1 LSB of the reference's own FFT path (oracle/_ref) — both truncate a value that lies
    within rounding error of the same integer;
  * differences from the exact product do occur (the conversion really truncates);
  * complete gates decrypt to the truth table.
Prints one JSON line; exit code 0 = all checks passed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle.pyoracle import Oracle, Ref, have_ref  # noqa: E402


def wrap32(d):
    d = np.asarray(d, dtype=np.int64)
    return (d + 2 ** 31) % 2 ** 32 - 2 ** 31


def main():
    pkg = ge.load_package()
    assert pkg.lib_path().endswith("_trunc.so"), pkg.lib_path()
    assert pkg.lib().tfhe_b200_conversion_mode() == 1
    o = Oracle()
    keys = o.keygen(42)
    eng = pkg.Engine(device=0)
    eng.load_keys(keys.bk, keys.ks)
    rng = np.random.default_rng(91)
    acc = rng.integers(-2 ** 31, 2 ** 31, size=(6, 2, 1024), dtype=np.int64).astype(np.int32)
    ref = None
    if have_ref():
        ref = Ref()
        ref.import_keys(keys)
    res = {"max_vs_exact": 0, "differing_words": 0, "words": 0, "max_vs_ref": None}
    for bk_index in (0, 3, 499):
        got = eng.extern_mul(eng.to_device(acc).clone(), bk_index).cpu().numpy()
        for i in range(acc.shape[0]):
            d = wrap32(got[i].astype(np.int64) - o.extern_mul_exact(keys.bk[bk_index], acc[i]).astype(np.int64))
            res["max_vs_exact"] = max(res["max_vs_exact"], int(np.abs(d).max()))
            res["differing_words"] += int((d != 0).sum())
            res["words"] += d.size
            if ref is not None:
                dr = wrap32(got[i].astype(np.int64) - ref.extern_mul(bk_index, acc[i]).astype(np.int64))
                res["max_vs_ref"] = max(res["max_vs_ref"] or 0, int(np.abs(dr).max()))
    # MuxRotate steps (rotation + external product + add), teacher forced
    cases = [(0, 1), (2, 2047), (5, 1024), (9, 777)]
    n_iter = max(i for i, _ in cases) + 1
    bara = np.zeros((len(cases), n_iter), np.int32)
    for r, (i, a) in enumerate(cases):
        bara[r, i] = a
    got = eng.blind_rotate(eng.to_device(acc[:len(cases)]).clone(), eng.to_device(bara)).cpu().numpy()
    for r, (i, a) in enumerate(cases):
        expect = o.blind_rotate_exact(keys.bk, acc[r], bara[r])
        d = wrap32(got[r].astype(np.int64) - expect.astype(np.int64))
        res["max_vs_exact"] = max(res["max_vs_exact"], int(np.abs(d).max()))
    # complete gates
    r2 = o.rng(5)
    ba, bb = rng.integers(0, 2, 64).astype(np.int32), rng.integers(0, 2, 64).astype(np.int32)
    ca, cb = o.encrypt_bits(keys, r2, ba), o.encrypt_bits(keys, r2, bb)
    out = eng.gate("NAND", eng.to_device(ca), eng.to_device(cb)).cpu().numpy()
    res["gates_ok"] = bool(np.array_equal(o.decrypt_bits(keys, out), 1 - (ba & bb)))
    eng.close()
    ok = (res["max_vs_exact"] <= 1 and res["differing_words"] > 0 and res["gates_ok"]
          and (res["max_vs_ref"] is None or res["max_vs_ref"] <= 1))
    res["ok"] = bool(ok)
    print(json.dumps(res))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
