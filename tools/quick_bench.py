"""Quick device-resident timing of gate batches (development aid, not the bench contract)."""
import sys, os, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle.pyoracle import Oracle

pkg = ge.load_package()
o = Oracle()
keys = o.keygen(42)
eng = pkg.Engine(device=0)
t = time.time(); eng.load_keys(keys.bk, keys.ks); torch.cuda.synchronize(); print("load_keys s", time.time() - t)
counts = [int(x) for x in sys.argv[1:]] or [1, 4, 592, 4736, 16384]
rng = np.random.default_rng(0)
for count in counts:
    bits_a = rng.integers(0, 2, count).astype(np.int32); bits_b = rng.integers(0, 2, count).astype(np.int32)
    # valid encryptions, vectorised: a uniform, b = <a,s> + mu + e
    def enc(bits):
        a = rng.integers(-2**31, 2**31, size=(count, 500), dtype=np.int64).astype(np.int32)
        e = np.rint(rng.normal(0, 2.44e-5 * 2**32, count)).astype(np.int64)
        mu = np.where(bits == 1, 2**29, -2**29)
        b = (a.astype(np.int64) * keys.lwe_key.astype(np.int64)).sum(1) + mu + e
        return np.concatenate([a, ((b + 2**31) % 2**32 - 2**31).astype(np.int32)[:, None]], axis=1)
    ca, cb = eng.to_device(enc(bits_a)), eng.to_device(enc(bits_b))
    out = eng.empty(count)
    for _ in range(2): eng.gate("NAND", ca, cb, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): eng.gate("NAND", ca, cb, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    dec = o.decrypt_bits(keys, out.cpu().numpy())
    ok = np.array_equal(dec, 1 - (bits_a & bits_b))
    print("count %6d  %.3f ms/batch  %.1f gates/s  bits_ok=%s" % (count, ms, count / ms * 1e3, ok), flush=True)
