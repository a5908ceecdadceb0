"""Times both blind-rotation schedules on several batch sizes (development aid)."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
sk = pkg.keygen(42)
eng = pkg.Engine(device=0); eng.load_keys(sk.bk, sk.ks)
counts = [int(x) for x in sys.argv[1:]] or [1, 4, 64, 592, 1184, 4736, 65536]
for count in counts:
    bits_a = np.random.default_rng(count).integers(0, 2, count).astype(np.int32)
    bits_b = np.random.default_rng(count + 1).integers(0, 2, count).astype(np.int32)
    ca = eng.to_device(pkg.encrypt_bits(sk, bits_a, 1)); cb = eng.to_device(pkg.encrypt_bits(sk, bits_b, 2))
    out = eng.empty(count)
    res = []
    for sched in (1, 2):
        eng.set_schedule(sched)
        for _ in range(2): eng.gate("NAND", ca, cb, out=out)
        torch.cuda.synchronize()
        reps = 3 if count > 10000 else 10
        eng.set_timing(True)
        for _ in range(reps): eng.gate("NAND", ca, cb, out=out)
        torch.cuda.synchronize()
        br, ks, n = eng.get_timing(); eng.set_timing(False)
        ok = np.array_equal(pkg.decrypt_bits(sk, out.cpu().numpy()), 1 - (bits_a & bits_b))
        res.append("sched %d: br %.3f ms ks %.3f ms ok=%s" % (sched, br / n, ks / n, ok))
    print("count %6d | %s | %s" % (count, res[0], res[1]), flush=True)
