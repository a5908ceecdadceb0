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


def key_shapes(p, with_secret_key=False):
    kpl = (p.k + 1) * p.l
    shapes = {
        "bk": (p.n, kpl, p.k + 1, p.N),
        "ks": (p.N * p.k, p.ks_t, 1 << p.ks_basebit, p.n + 1),
    }
    if with_secret_key:
        shapes["lwe_key"] = (p.n,)
    return shapes


def broadcast_cloud_keys(params, sk, device, src=0, with_secret_key=False):
    """Rank `src` passes its SecretKeys (others pass None); every rank gets int32 tensors
    bk [n][kpl][k+1][N] and ks [N][t][base][n+1] on `device`: the CLOUD keys only.
    with_secret_key=True additionally ships the LWE secret key ("lwe_key") — an explicit opt-in
    for benchmarks and tests that verify every rank's outputs by decrypting them; a deployment
    never does this."""
    import torch
    import torch.distributed as dist

    shapes = key_shapes(params, with_secret_key)
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


def gather_rows(local, sizes):
    """All-gather of per-rank row blocks with the given sizes (rows per rank)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return local
    maxlen = max(sizes)
    pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)


class ShardedMatmul:
    """C = A (rows x inner) * B (inner x cols) of nbits-bit integers over the ranks (BASELINE config 5;
    SURVEY.md 8e): the rows of C are split contiguously, every rank holds the whole encrypted A and B
    (2*256*8 samples = 8 MB for the 16x16 case) and runs the plan of its own rows
    (tfhe_b200_circuit_matmul_ex); the only exchange is the final gather of the result rows.
    engine None: plaintext simulation of the same plans (CPU tests)."""

    def __init__(self, pkg, engine, rows, inner, cols, nbits, adder=0, world=None, rank=None):
        import torch.distributed as dist

        multi = dist.is_available() and dist.is_initialized()
        self.world = world if world is not None else (dist.get_world_size() if multi else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if multi else 0)
        self.rows, self.inner, self.cols, self.nbits = rows, inner, cols, nbits
        self.bounds = [shard_bounds(rows, self.world, r) for r in range(self.world)]
        self.lo, self.hi = self.bounds[self.rank]
        self.circ = pkg.Circuit(engine, "matmul_ex", self.hi - self.lo, inner, cols, nbits, adder) \
            if self.hi > self.lo else None
        self.sizes = [(hi - lo) * cols * nbits for lo, hi in self.bounds]

    def local_rows_of_A(self, A):
        """A: [rows*inner*nbits, ...] (samples or plaintext bits) -> this rank's rows."""
        per = self.inner * self.nbits
        return A[self.lo * per: self.hi * per]

    def run(self, A_enc, B_enc, gather=True):
        import torch

        if self.circ is not None:
            local = self.circ.run(self.local_rows_of_A(A_enc).contiguous(), B_enc)
        else:
            local = torch.empty((0, A_enc.shape[1]), dtype=A_enc.dtype, device=A_enc.device)
        return gather_rows(local, self.sizes) if gather else local

    def simulate(self, A_bits, B_bits):
        import torch

        A_bits, B_bits = np.asarray(A_bits).reshape(-1), np.asarray(B_bits).reshape(-1)
        local = self.circ.simulate(self.local_rows_of_A(A_bits), B_bits) if self.circ is not None \
            else np.zeros(0, np.int32)
        return gather_rows(torch.from_numpy(np.ascontiguousarray(local, np.int32)).reshape(-1, 1), self.sizes) \
            .reshape(-1).numpy()
