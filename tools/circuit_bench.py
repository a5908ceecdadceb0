"""Latency of the Cipher-level circuits (BASELINE.json configs 1, 2, 4, 5) on one B200:
single bootstrapped gate, 16-bit addition (both reference schedules), 32-bit multiplication,
16x16 matrix multiply of 8-bit integers.  Results are decrypted and checked.  Prints JSON."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
sk = pkg.keygen(2026)
eng = pkg.Engine(device=0)
eng.load_keys(sk.bk, sk.ks)


def to_bits(vals, nbits):
    vals = np.asarray(vals, dtype=np.int64).reshape(-1)
    return ((vals[:, None] >> np.arange(nbits)) & 1).astype(np.int32)


def enc(vals, nbits, seed):
    return eng.to_device(pkg.encrypt_bits(sk, to_bits(vals, nbits).reshape(-1), seed))


def dec(t, nbits):
    bits = pkg.decrypt_bits(sk, t.cpu().numpy()).reshape(-1, nbits).astype(np.int64)
    return (bits << np.arange(nbits)).sum(-1)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


res = {}
# config 1: single bootsNAND
ca, cb = enc([1], 1, 1), enc([1], 1, 2)
ms, out = timed(lambda: eng.gate("NAND", ca, cb), 20)
assert dec(out, 1)[0] == 0
res["single_gate_ms"] = ms
# config 2: 16-bit addition
a, b = np.array([12345]), np.array([(-6789) & 0xFFFF])
for mode, name in ((0, "add16_bitwise"), (1, "add16_numberwise"), (2, "add16_prefix")):
    c = pkg.Circuit(eng, "add", 16, 1, mode)
    da, db = enc(a, 16, 3), enc(b, 16, 4)
    ms, out = timed(lambda: c.run(da, db), 3)
    assert dec(out, 16)[0] == (a[0] + b[0]) & 0xFFFF
    res[name] = {"ms": ms, "levels": c.levels, "gates": c.gates}
# config 4: 32-bit multiplication
da, db = enc([40000], 32, 5), enc([50000], 32, 6)
for adder, name in ((0, "mul32"), (1, "mul32_prefix")):
    c = pkg.Circuit(eng, "mul_ex", 32, 1, adder)
    ms, out = timed(lambda: c.run(da, db), 1)
    assert dec(out, 32)[0] == (40000 * 50000) & 0xFFFFFFFF
    res[name] = {"ms": ms, "levels": c.levels, "gates": c.gates}
# config 5: 16x16 matrix multiply of 8-bit integers (one GPU here; the 256 output elements shard over GPUs)
n = int(os.environ.get("MATMUL_N", "16"))
rng = np.random.default_rng(1)
A, B = rng.integers(-8, 8, (n, n)), rng.integers(-8, 8, (n, n))
da, db = enc(A.reshape(-1) & 0xFF, 8, 7), enc(B.reshape(-1) & 0xFF, 8, 8)
for adder, name in ((0, "matmul%dx%d_8bit" % (n, n)), (1, "matmul%dx%d_8bit_prefix" % (n, n))):
    t0 = time.time(); c = pkg.Circuit(eng, "matmul_ex", n, n, n, 8, adder); plan_s = time.time() - t0
    ms, out = timed(lambda: c.run(da, db), 1)
    assert np.array_equal(dec(out, 8).reshape(n, n), (A @ B) & 0xFF)
    res[name] = {"ms": ms, "levels": c.levels, "gates": c.gates, "plan_build_s": plan_s,
                 "gates_per_s": c.gates / ms * 1e3}
    c.close()
print(json.dumps(res))
