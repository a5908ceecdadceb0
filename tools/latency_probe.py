"""Development aid: latency of small gate batches, split into blind rotation and key switch
(tfhe_b200_set_timing), warm GPU, many repetitions.  usage: latency_probe.py [counts...]"""
import os, sys
import threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
sk = pkg.keygen(1)
eng = pkg.Engine(device=0); eng.load_keys(sk.bk, sk.ks)
rng = np.random.default_rng(0)
counts = [int(x) for x in sys.argv[1:]] or [1, 2, 16, 148, 296, 444, 592]
try:
    import pynvml
    pynvml.nvmlInit(); _h = pynvml.nvmlDeviceGetHandleByIndex(0)
    def sm_clock(): return pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
    def power(): return pynvml.nvmlDeviceGetPowerUsage(_h) / 1000.0
except Exception:
    def sm_clock(): return -1
    def power(): return -1
# warm the clocks with a large batch
big = 4736
wa = eng.to_device(pkg.encrypt_bits(sk, rng.integers(0, 2, big).astype(np.int32), 1))
for _ in range(3): eng.gate("NAND", wa, wa)
torch.cuda.synchronize()
for count in counts:
    ba, bb = rng.integers(0, 2, count).astype(np.int32), rng.integers(0, 2, count).astype(np.int32)
    ca, cb = eng.to_device(pkg.encrypt_bits(sk, ba, 2)), eng.to_device(pkg.encrypt_bits(sk, bb, 3))
    out = eng.empty(count)
    for _ in range(5): eng.gate("NAND", ca, cb, out=out)
    torch.cuda.synchronize()
    reps = 120
    clocks, stop = [], [False]
    def sample():
        while not stop[0]:
            clocks.append((sm_clock(), power())); time.sleep(0.02)
    th = threading.Thread(target=sample); th.start()
    eng.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): eng.gate("NAND", ca, cb, out=out)
    e1.record(); torch.cuda.synchronize()
    br, ks, calls = eng.get_timing(); eng.set_timing(False)
    stop[0] = True; th.join()
    ok = np.array_equal(pkg.decrypt_bits(sk, out.cpu().numpy()), 1 - (ba & bb))
    cl = sorted(c for c, _ in clocks) or [-1]
    print("count %4d: %.3f ms per batch  (blind rotation %.3f, key switch %.3f)  ok=%s  sm clock MHz min/med/max %d/%d/%d  power %.0f W" % (
        count, e0.elapsed_time(e1) / reps, br / calls, ks / calls, ok, cl[0], cl[len(cl) // 2], cl[-1],
        max([p for _, p in clocks] or [-1])), flush=True)
