"""Profiling aid: `count` blind rotations of `n_iter` dense iterations (every bara != 0)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
sk = pkg.keygen(1)
eng = pkg.Engine(device=0); eng.load_keys(sk.bk, sk.ks)
rng = np.random.default_rng(12)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 148
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
acc = rng.integers(-2**31, 2**31, size=(count, 2, 1024), dtype=np.int64).astype(np.int32)
bara = rng.integers(1, 2048, size=(count, n_iter)).astype(np.int32)
d_acc, d_bara = eng.to_device(acc), eng.to_device(bara)
for _ in range(2):
    got = eng.blind_rotate(d_acc.clone(), d_bara)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); got = eng.blind_rotate(d_acc.clone(), d_bara); e1.record(); torch.cuda.synchronize()
print("count %d n_iter %d: %.3f ms, %.3f us per iteration-wave" % (count, n_iter, e0.elapsed_time(e1),
      1e3 * e0.elapsed_time(e1) / n_iter / max(1, -(-count // (4 * int(eng.sm_count))))))
