"""BASELINE config 5 on N GPUs: encrypted 16x16 matrix multiply of 8-bit integers, rows of C
sharded over the ranks (cpu-gpu-tfhe_b200/dist.py ShardedMatmul), keys and the encrypted operands
broadcast once over NCCL.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port P tools/matmul_multi.py   (or plain python for N = 1).
Rank 0 decrypts and checks the result and prints one JSON line (time = max over ranks)."""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from importlib import import_module

pkg = ge.load_package()
D = import_module("cpu_gpu_tfhe_b200.dist")
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(os.environ.get("MATMUL_N", "16"))
nbits = 8
adder = int(os.environ.get("ADDER", "0"))
p = pkg.default_params()
sk = pkg.keygen(2026) if rank == 0 else None
t0 = time.time()
keys = D.broadcast_cloud_keys(p, sk, dev)
eng = pkg.Engine(device=local)
eng.load_keys_device(keys["bk"], keys["ks"])
torch.cuda.synchronize()
key_s = time.time() - t0
rng = np.random.default_rng(1)
A, B = rng.integers(-8, 8, (n, n)), rng.integers(-8, 8, (n, n))
bits = lambda v: ((np.asarray(v).reshape(-1)[:, None] % 2 ** nbits >> np.arange(nbits)) & 1).astype(np.int32)
shape = (2 * n * n * nbits, p.n + 1)
if rank == 0:
    enc = torch.from_numpy(np.concatenate([pkg.encrypt_bits(sk, bits(A).reshape(-1), 7),
                                           pkg.encrypt_bits(sk, bits(B).reshape(-1), 8)])).to(dev)
else:
    enc = torch.empty(shape, dtype=torch.int32, device=dev)
if world > 1:
    dist.broadcast(enc, 0)
eA, eB = enc[: n * n * nbits].contiguous(), enc[n * n * nbits:].contiguous()
sm = D.ShardedMatmul(pkg, eng, n, n, n, nbits, adder)
sm.run(eA, eB, gather=False)  # warm-up (uploads the plan)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = sm.run(eA, eB)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    got = pkg.decrypt_bits(sk, out.cpu().numpy()).reshape(-1, nbits).astype(np.int64)
    C = (got << np.arange(nbits)).sum(-1).reshape(n, n)
    ok = bool(np.array_equal(C, (A @ B) % 2 ** nbits))
    gates = sm.circ.gates * world if n % world == 0 else None
    print(json.dumps({"workload": "%dx%d matmul of %d-bit integers" % (n, n, nbits), "n_gpus": world,
                      "adder": "prefix" if adder else "ripple", "ms": float(ms.item()), "correct": ok,
                      "levels": sm.circ.levels, "gates_total": gates, "key_broadcast_s": key_s}))
    assert ok
if world > 1:
    dist.destroy_process_group()
