"""Development aid: times the key-switch kernel alone on random extracted samples and checks it
against the previous build (TFHE_B200_LIB_REF) if given, bit for bit."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
sk = pkg.keygen(3)
eng = pkg.Engine(device=0); eng.load_keys(sk.bk, sk.ks)
for count in [int(x) for x in sys.argv[1:]] or [1, 100, 4736, 65536]:
    g = torch.Generator(device="cuda").manual_seed(count)
    u = torch.randint(-2**31, 2**31 - 1, (count, 1025), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    out = eng.keyswitch(u); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = eng.keyswitch(u)
    e1.record(); torch.cuda.synchronize()
    print("count %6d  %.3f ms  checksum %d" % (count, e0.elapsed_time(e1) / 3, int(out.to(torch.int64).sum().item())), flush=True)
