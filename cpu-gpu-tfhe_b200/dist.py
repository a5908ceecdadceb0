"""Multi-GPU plumbing: one process per GPU, torch.distributed for the one-time key broadcast.

Bootstrapped gates are independent: a batch is split contiguously over the
ranks, every rank holds the full (read-only) key material, and nothing is exchanged on the
data path.  The only collective is the broadcast of the cloud keys from rank 0 (NCCL over
NVLink on GPUs; gloo in the CPU tests) and, if the caller wants the results in one place, a
gather of the outputs.
"""
import numpy as np


def shard_bounds(total, world, rank):
    """Contiguous split of `total` gates: rank r gets [lo, hi); sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def key_shapes(p):
    kpl = (p.k + 1) * p.l
    return {
        "bk": (p.n, kpl, p.k + 1, p.N),
        "ks": (p.N * p.k, p.ks_t, 1 << p.ks_basebit, p.n + 1),
        "lwe_key": (p.n,),
    }


def broadcast_cloud_keys(params, sk, device, src=0):
    """Rank `src` passes its SecretKeys (others pass None); every rank gets int32 tensors
    bk [n][kpl][k+1][N] and ks [N][t][base][n+1] on `device` (plus the LWE key, which the
    benchmark uses to verify outputs — a real deployment would not ship it)."""
    import torch
    import torch.distributed as dist

    shapes = key_shapes(params)
    out = {}
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else src
    for name, shape in shapes.items():
        if rank == src:
            t = torch.from_numpy(np.ascontiguousarray(getattr(sk, name), dtype=np.int32)).to(device)
        else:
            t = torch.empty(shape, dtype=torch.int32, device=device)
        if multi:
            dist.broadcast(t, src)
        out[name] = t
    return out


def gather_outputs(local, total, device=None):
    """All-gather of the per-rank output shards (rows of n+1 words) into [total, n+1]."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    width = local.shape[1]
    maxlen = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxlen, width), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
