"""Development aid: per-phase cycle counts of one warp of the blind-rotation kernel.
Needs a library built with -DTFHE_B200_PHASE_TIMING=1 (tools/build_variant.py):
    python tools/build_variant.py pt -DTFHE_B200_PHASE_TIMING=1
    TFHE_B200_LIB=build/variants/pt/libtfhe_b200.so python tools/phase_timing.py 148 296 592
counts <= 148 put one ciphertext on each SM (a lone warp pair), 592 four."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
lib = pkg.lib()
sk = pkg.keygen(1)
eng = pkg.Engine(device=0); eng.load_keys(sk.bk, sk.ks)
rng = np.random.default_rng(12)
names = ["f1 (decomp, pass 1, store)", "f2_fft (2 rows)", "mac keep (+wait)", "mac give (+wait)", "xchg_store",
         "pair barrier", "xchg_load + inv16", "i2_inner", "i2 shuffles", "i2_final", "-", "-", "-", "-", "-", "-"]
if os.environ.get("QUAD_NAMES"):
    names = ["f1 (decomp, pass 1, store)", "barrier 1", "pass 2 in place (warp cq) + pair barrier", "load rows + mac (+wait)",
             "inverse stages 3, 2 + store", "pair barrier", "inverse stages 1, 0 in place", "barrier 2", "i2_local",
             "shuffles + final", "pair barrier (i2)", "-", "-", "-", "-", "-"]
n_iter = 100
buf = (ctypes.c_longlong * 16)()
for count in [int(x) for x in sys.argv[1:]] or [148, 592]:
    acc = rng.integers(-2**31, 2**31, size=(count, 2, 1024), dtype=np.int64).astype(np.int32)
    bara = rng.integers(1, 2048, size=(count, n_iter)).astype(np.int32)
    d_acc, d_bara = eng.to_device(acc), eng.to_device(bara)
    eng.blind_rotate(d_acc.clone(), d_bara); torch.cuda.synchronize()
    lib.tfhe_b200_debug_phase_cycles(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.blind_rotate(d_acc.clone(), d_bara); e1.record(); torch.cuda.synchronize()
    lib.tfhe_b200_debug_phase_cycles(buf, 1)
    tot = sum(buf[i] for i in range(12))
    print("count %d: %.3f ms, %.0f cycles per iteration (sum of phases), wall %.2f us/iter" % (
        count, e0.elapsed_time(e1), tot / n_iter, 1e3 * e0.elapsed_time(e1) / n_iter))
    for i, nm in enumerate(names):  # the ring rows are a breakdown of the two mac rows
        if nm != "-":
            print("   %-28s %8.0f  %5.1f%%" % (nm, buf[i] / n_iter, 100.0 * buf[i] / tot))
