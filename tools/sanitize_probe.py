"""Development aid: a few small launches of every hot kernel for compute-sanitizer
(memcheck / racecheck / synccheck).  usage: compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
p = pkg.default_params()
sk = pkg.keygen(3)
eng = pkg.Engine(device=0); eng.load_keys(sk.bk, sk.ks)
rng = np.random.default_rng(1)
def rot(count, n_iter, skip=0.2):
    acc = rng.integers(-2**31, 2**31, size=(count, 2, 1024), dtype=np.int64).astype(np.int32)
    bara = rng.integers(1, 2048, size=(count, n_iter)).astype(np.int32)
    bara[rng.random((count, n_iter)) < skip] = 0
    out = eng.blind_rotate(eng.to_device(acc).clone(), eng.to_device(bara)); torch.cuda.synchronize()
    return out
rot(3, 6)                       # latency kernel (one ciphertext per CTA)
rot(eng.sm_count + 9, 4)        # small-batch instantiation of the throughput kernel (helper slot)
rot(4 * eng.sm_count + 7, 3)    # full-batch throughput kernel, ragged tail
bits = rng.integers(0, 2, 40).astype(np.int32)
c = eng.to_device(pkg.encrypt_bits(sk, bits, 5))
out = eng.gate("NAND", c, c); m = eng.mux(c, c, c); torch.cuda.synchronize()
assert np.array_equal(pkg.decrypt_bits(sk, out.cpu().numpy()), 1 - bits)
big = eng.to_device(pkg.encrypt_bits(sk, rng.integers(0, 2, 2304).astype(np.int32), 6))
ks = eng.gate("AND", big, big); torch.cuda.synchronize()   # tensor-core key switch
print("sanitize probe done")
