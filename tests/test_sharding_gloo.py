"""N > 1 path on CPU: two gloo ranks broadcast the cloud keys and shard a gate batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    pkg = ge.load_package()
    from importlib import import_module

    d = import_module("cpu_gpu_tfhe_b200.dist")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = pkg.default_params()
    p.n = 12  # small LWE dimension keeps key generation and the broadcast cheap
    sk = pkg.keygen(5, p) if rank == 0 else None
    keys = d.broadcast_cloud_keys(p, sk, torch.device("cpu"))
    lo, hi = d.shard_bounds(total, world, rank)
    # every rank "computes" its shard: here the identity on a deterministic batch
    full = torch.arange(total * (p.n + 1), dtype=torch.int32).reshape(total, p.n + 1)
    gathered = d.gather_outputs(full[lo:hi].clone(), total)
    q.put((rank, lo, hi, int(keys["bk"].sum()), int(keys["ks"].sum()), bool(torch.equal(gathered, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 65536, 7])
def test_two_rank_key_broadcast_and_sharding(total):
    world, port = 2, 29500 + os.getpid() % 2000 + total % 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, bk0, ks0, ok0), (r1, lo1, hi1, bk1, ks1, ok1) = res
    assert (lo0, hi1) == (0, total) and hi0 == lo1 and abs((hi0 - lo0) - (hi1 - lo1)) <= 1
    assert bk0 == bk1 and ks0 == ks1 and bk0 != 0  # both ranks hold rank 0's keys
    assert ok0 and ok1


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.load_package()
    from importlib import import_module

    d = import_module("cpu_gpu_tfhe_b200.dist")
    for total in (0, 1, 7, 65536, 1000003):
        for world in (1, 2, 4, 8):
            b = [d.shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _matmul_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    pkg = ge.load_package()
    from importlib import import_module

    d = import_module("cpu_gpu_tfhe_b200.dist")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, inner, cols, nbits = 5, 3, 4, 6
    rng = np.random.default_rng(1)
    A, B = rng.integers(-8, 8, (rows, inner)), rng.integers(-8, 8, (inner, cols))
    bits = lambda v: ((np.asarray(v).reshape(-1)[:, None] % 2 ** nbits >> np.arange(nbits)) & 1).astype(np.int32)
    sm = d.ShardedMatmul(pkg, None, rows, inner, cols, nbits, adder=rank % 2)  # ranks may even differ in adder
    out = sm.simulate(bits(A), bits(B)).reshape(-1, nbits)
    got = (out.astype(np.int64) << np.arange(nbits)).sum(-1).reshape(rows, cols)
    q.put((rank, sm.lo, sm.hi, bool(np.array_equal(got, (A @ B) % 2 ** nbits))))
    dist.destroy_process_group()


def test_two_rank_sharded_matrix_multiply_plan():
    """BASELINE config 5 sharding (rows of C over the ranks, result gather) on plaintext plans."""
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_matmul_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    assert all(r[3] for r in res)
