#!/usr/bin/env python3
"""Headline benchmark: bootstrapped gates/s on B200 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle/_ref)

A step = one batch of independent bootstrapped NAND gates per GPU (BASELINE.json
configs[2]: 65536 gates; each rank processes its own 65536, scaling "weak": the
gates are independent and nothing is exchanged on the data path; keys are
generated on rank 0 and broadcast once over NCCL before timing).

Prints ONE JSON line (rank 0): metric/value/unit, ms_per_step, e2e (host buffers,
copies inside the timed region, through the C ABI), roofline of the dominant kernel
(fp64 blind rotation, live CUDA-event kernel time / measured fp64 peak), cpu_baseline
(reference host path on this box's cores), clocks, gpu_launches.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bootstrapped_gates_per_sec"
UNIT = "gates/s"
FLOP_PER_BOOTSTRAP = 500 * (6 * 26112 + 32768)  # 94.72 MFLOP (algorithmic, folded FFT)
BATCH = 65536


# ----------------------------------------------------------------- CPU side ---

def _load_keys(path):
    from oracle.pyoracle import Keys, Oracle

    o = Oracle()
    z = {k: np.load(os.path.join(path, k + ".npy"), mmap_mode="r") for k in ("lwe", "tlwe", "bk", "ks", "ca", "cb")}
    return o, Keys(o.params, np.array(z["lwe"]), np.array(z["tlwe"]), z["bk"], z["ks"]), np.array(z["ca"]), np.array(z["cb"])


def _ref_worker(args):
    """One process per core: the reference's own host path (oracle/_ref) on `count` NAND gates."""
    path, count = args
    from oracle.pyoracle import Ref

    o, keys, ca, cb = _load_keys(path)
    r = Ref().import_keys(keys)
    secs, last = r.time_nand(ca, cb, count)
    return secs, int(r.phase(last) > 0)


def _port_worker(args):
    path, count = args
    from oracle.pyoracle import FFT_FOLDED

    o, keys, ca, cb = _load_keys(path)
    cx = o.ctx(keys, FFT_FOLDED)
    t0 = time.perf_counter()
    for g in range(count):
        cx.gate("NAND", ca[g % len(ca)], cb[g % len(cb)])
    return time.perf_counter() - t0, 0


_KEY_DIR = None


def _dump_keys(sk, pkg):
    """Keys + 16 input pairs written once to a temp dir; workers mmap them."""
    global _KEY_DIR
    if _KEY_DIR is None:
        import tempfile

        _KEY_DIR = tempfile.mkdtemp(prefix="tfhe_b200_bench_")
        bits_a, bits_b = np.arange(16) % 2, (np.arange(16) // 2) % 2
        for name, arr in (("lwe", sk.lwe_key), ("tlwe", sk.tlwe_key), ("bk", sk.bk), ("ks", sk.ks),
                          ("ca", pkg.encrypt_bits(sk, bits_a, 11)), ("cb", pkg.encrypt_bits(sk, bits_b, 12))):
            np.save(os.path.join(_KEY_DIR, name + ".npy"), np.ascontiguousarray(arr))
    return _KEY_DIR


def cpu_reference_run(sk, pkg, gates_per_core, cores=None):
    """Times the reference CPU implementation on `cores` processes x gates_per_core NANDs
    (the reference host path is not re-entrant -- global FFT scratch, lagrangehalfc_impl.cu:4 --
    so parallelism is one process per core).  Returns dict(value gates/s, cores, kind, sample)."""
    from oracle.pyoracle import have_ref

    cores = cores or os.cpu_count() or 1
    kind = "reference" if have_ref() else "port"
    path = _dump_keys(sk, pkg)
    worker = _ref_worker if kind == "reference" else _port_worker
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(worker, [(path, gates_per_core)] * cores, chunksize=1)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)  # slowest worker's gate loop (excludes process start / key import)
    total = cores * gates_per_core
    return {
        "value": total / busy, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": "%d bootsNAND per core on %d cores (%s host path, FFTW-shim FFT), %.1f s gate time, %.1f s wall"
                  % (gates_per_core, cores, "reference oracle/_ref" if kind == "reference" else "C port", busy, wall),
        "ms_per_gate_per_core": 1e3 * busy / gates_per_core,
    }


def _port_rate_worker(args):
    """One process per core: the oracle's folded-FFT port (N/2-point twisted transform, the
    formulation spqlios uses) on `count` NAND gates."""
    return _port_worker(args)


def cpu_port_folded_rate(sk, pkg, gates_per_core, cores):
    path = _dump_keys(sk, pkg)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_port_rate_worker, [(path, gates_per_core)] * cores, chunksize=1)
    busy = max(r[0] for r in res)
    return {"value": cores * gates_per_core / busy, "unit": UNIT, "cores": cores,
            "ms_per_gate_per_core": 1e3 * busy / gates_per_core,
            "what": "oracle C port, folded N/2-point FFT (the spqlios formulation, plain C, no AVX assembly)"}


def cpu_circuit_baseline(sk, pkg, cores):
    """The CPU side of BASELINE configs 2, 4, 5: the CPU reference's own Cipher schedules
    (cpuParallel/Cipher.cpp operator+ :348-375, operator* with its OpenMP reduction :83-112, the
    matrix-multiply loop body cpuParallel/cloud.cpp:390-408) executed by the oracle's C restatement
    in FFT_REF mode (bit-identical to oracle/_ref, tests/test_oracle_vs_ref.py; the reference's own
    objects are not re-entrant, so OpenMP schedules cannot run on them).  Bounded samples: a 16-bit
    add, an 8-bit multiply with the OpenMP reduction, and one loop body per core; the 32-bit multiply
    and the 16x16 matrix are extrapolated with the stated formulas."""
    from oracle.pyoracle import FFT_REF, Keys, Oracle

    o = Oracle()
    keys = Keys(o.params, sk.lwe_key, sk.tlwe_key, sk.bk, sk.ks)
    cx = o.ctx(keys, FFT_REF)

    def bits(v, n):
        return ((np.asarray([v], dtype=np.int64)[:, None] >> np.arange(n)) & 1).astype(np.int32).reshape(-1)

    def enc(v, n, seed):
        return pkg.encrypt_bits(sk, bits(v, n), seed)

    def dec(c):
        b = pkg.decrypt_bits(sk, c).astype(np.int64)
        return int((b << np.arange(b.size)).sum())

    res, ok = {}, True
    a, b = 12345, (-6789) & 0xFFFF
    th = min(8, cores)
    ea, eb, ma, mb = enc(a, 16, 21), enc(b, 16, 22), enc(201, 8, 23), enc(57, 8, 24)
    box = {}

    def run_add():
        t0 = time.perf_counter()
        box["add"] = cx.add(ea, eb)
        res["add16_ms"] = 1e3 * (time.perf_counter() - t0)

    def run_mul():
        t0 = time.perf_counter()
        box["mul"] = cx.cipher_mul(ma, mb, th)
        res["mul8_omp_ms"] = 1e3 * (time.perf_counter() - t0)

    if cores >= th + 2:  # enough cores: the sequential add runs beside the 8-thread multiply
        ta = threading.Thread(target=run_add)
        ta.start()
        run_mul()
        ta.join()
    else:
        run_add()
        run_mul()
    res["mul8_omp_threads"] = th
    ok = ok and dec(box["add"]) == (a + b) & 0xFFFF and dec(box["mul"]) == 201 * 57
    # critical path of operator* with T threads: ceil(n/T) * (n ANDs + 2n*5 adder gates) + T * 2n*5 (combine)
    path = lambda n, T: -(-n // T) * (n + 10 * n) + T * 10 * n
    t32 = min(32, cores)
    res["mul32_omp_estimate_ms"] = res["mul8_omp_ms"] * path(32, t32) / path(8, th)
    res["mul32_omp_estimate_how"] = ("mul8 time x critical-path gate ratio %d/%d (%d threads; ceil(n/T)*11n + T*10n gates)"
                                     % (path(32, t32), path(8, th), t32))
    units = cores
    A = np.stack([enc(3 + u % 5, 8, 30 + u) for u in range(units)])
    B = np.stack([enc(7 + u % 3, 8, 130 + u) for u in range(units)])
    C = np.stack([enc(100 + u, 16, 230 + u) for u in range(units)])
    t0 = time.perf_counter()
    out = cx.matmul_units(A, B, C, cores)
    unit_s = time.perf_counter() - t0
    ok = ok and all(dec(out[u]) == (100 + u + (3 + u % 5) * (7 + u % 3)) & 0xFFFF for u in range(units))
    res["matmul_unit_s"] = unit_s
    res["matmul16x16_8bit_estimate_s"] = -(-256 // cores) * 16 * unit_s
    res["matmul16x16_8bit_estimate_how"] = ("ceil(256 output elements / %d cores) x 16 loop bodies x %.1f s per body "
                                            "(8-bit operator* + 16-bit add, one body per core, all cores busy)"
                                            % (cores, unit_s))
    res["results_decrypt_ok"] = bool(ok)
    res["cores"] = cores
    res["kind"] = "port of cpuParallel/Cipher.cpp schedules on the oracle C restatement (FFT_REF mode = oracle/_ref arithmetic)"
    return res


# ------------------------------------------------------------- clock sampler ---

class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# -------------------------------------------------------------------- arms ---

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge

    pkg = ge.load_package()
    sk = pkg.keygen(2026)
    cores = os.cpu_count() or 1
    # size the per-step sample from a one-gate probe so that K+W steps stay within a few minutes
    probe = cpu_reference_run(sk, pkg, 2, cores)
    per_gate = probe["ms_per_gate_per_core"] / 1e3
    budget = 120.0 / max(1, args.steps + args.warmup)
    gpc = max(2, min(64, int(budget / per_gate)))
    vals = []
    for step in range(args.warmup + args.steps):
        r = cpu_reference_run(sk, pkg, gpc, cores)
        if step >= args.warmup:
            vals.append(r)
    gates = gpc * cores * len(vals)
    secs = sum(gpc * cores / v["value"] for v in vals)
    value = gates / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(vals),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "independent bootstrapped NAND gates, default TFHE gate params (n=500,N=1024,k=1,l=2)",
                   "gates_per_step": gpc * cores, "note": "bounded sample of the 65536-gate batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": vals[0]["kind"],
                         "sample": vals[0]["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def circuit_latencies(pkg, eng, sk):
    """BASELINE.json's latency metrics on one GPU (configs 1, 2, 4): single bootstrapped gate, 16-bit
    addition and 32-bit multiplication, reference schedules and the parallel-prefix ones; results are
    decrypted and checked.  Device time (CUDA events), plans built and warmed up outside the timing."""
    import torch

    def to_bits(v, nbits):
        return ((np.asarray([v], dtype=np.int64)[:, None] >> np.arange(nbits)) & 1).astype(np.int32).reshape(-1)

    def enc(v, nbits, seed):
        return eng.to_device(pkg.encrypt_bits(sk, to_bits(v, nbits), seed))

    def dec(t, nbits):
        bits = pkg.decrypt_bits(sk, t.cpu().numpy()).reshape(-1, nbits).astype(np.int64)
        return int((bits << np.arange(nbits)).sum(-1)[0])

    def timed(fn, reps, stream=None):
        with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream()):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                out = fn()
            e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    res = {}
    side = torch.cuda.Stream()  # plans replay a captured CUDA graph on a non-default stream
    torch.cuda.synchronize()
    ca, cb = enc(1, 1, 1), enc(1, 1, 2)
    for _ in range(5):
        eng.gate("NAND", ca, cb)
    runs = [timed(lambda: eng.gate("NAND", ca, cb), 20) for _ in range(5)]
    ms, out = min(r[0] for r in runs), runs[-1][1]
    res["single_gate_ms"] = ms
    res["single_gate_ms_runs"] = [r[0] for r in runs]
    ok = dec(out, 1) == 0
    a, b = 12345, (-6789) & 0xFFFF
    da, db = enc(a, 16, 3), enc(b, 16, 4)
    for mode, name in ((0, "add16_reference_bitwise"), (1, "add16_reference_numberwise"), (2, "add16_prefix")):
        c = pkg.Circuit(eng, "add", 16, 1, mode)
        ms, out = timed(lambda: c.run(da, db), 2, side)
        ok = ok and dec(out, 16) == (a + b) & 0xFFFF
        res[name] = {"ms": ms, "levels": c.levels, "gates": c.gates, "cuda_graph": c.used_graph}
        c.close()
    # eight independent 16-bit additions merged level by level (tfhe_b200_circuit_run_many)
    plans = [pkg.Circuit(eng, "add", 16, 1, 2) for _ in range(8)]
    ops = [[enc(1000 * i + 7, 16, 40 + i), enc(555 * i + 1, 16, 60 + i)] for i in range(8)]
    ms, outs = timed(lambda: pkg.Circuit.run_many(plans, ops), 2, side)
    ok = ok and all(dec(o, 16) == (1000 * i + 7 + 555 * i + 1) & 0xFFFF for i, o in enumerate(outs))
    res["add16_prefix_x8_merged"] = {"ms": ms, "ms_per_addition": ms / 8, "plans": 8}
    for c in plans:
        c.close()
    da, db = enc(40000, 32, 5), enc(50000, 32, 6)
    for adder, name in ((0, "mul32_reference"), (1, "mul32_prefix"), (2, "mul32_carry_save")):
        c = pkg.Circuit(eng, "mul_ex", 32, 1, adder)
        ms, out = timed(lambda: c.run(da, db), 1, side)
        ok = ok and dec(out, 32) == (40000 * 50000) & 0xFFFFFFFF
        res[name] = {"ms": ms, "levels": c.levels, "gates": c.gates, "cuda_graph": c.used_graph}
        c.close()
    res["results_decrypt_ok"] = bool(ok)
    return res


def matmul_config5(pkg, eng, tdist, sk, p, dev, world, rank, barrier):
    """BASELINE configs[4]: 16x16 matrix multiply of 8-bit integers on the launched GPUs
    (cpu-gpu-tfhe_b200/dist.py ShardedMatmul: rows of C sharded, every rank holds the encrypted A and
    B, the only exchange is the final gather).  Two schedules: a carry-save tree per element of C (the product's
    own) and the reference's ripple adders (BOOTS_matrixMultiplication, main.cu:2342).  Device time,
    max over ranks; rank 0 decrypts and checks."""
    import torch
    import torch.distributed as dist

    nmat, nbits = 16, 8
    rng = np.random.default_rng(1)
    A, B = rng.integers(-8, 8, (nmat, nmat)), rng.integers(-8, 8, (nmat, nmat))
    bits = lambda v: ((np.asarray(v).reshape(-1)[:, None] % 2 ** nbits >> np.arange(nbits)) & 1).astype(np.int32)
    rows = nmat * nmat * nbits
    if rank == 0:
        enc = torch.from_numpy(np.concatenate([pkg.encrypt_bits(sk, bits(A).reshape(-1), 7),
                                               pkg.encrypt_bits(sk, bits(B).reshape(-1), 8)])).to(dev)
    else:
        enc = torch.empty((2 * rows, p.n + 1), dtype=torch.int32, device=dev)
    if world > 1:
        dist.broadcast(enc, 0)
    eA, eB = enc[:rows].contiguous(), enc[rows:].contiguous()
    res = {}
    for adder, name, warm in ((2, "matmul16x16_8bit", True), (0, "matmul16x16_8bit_reference_schedule", False)):
        sm = tdist.ShardedMatmul(pkg, eng, nmat, nmat, nmat, nbits, adder)
        if warm:
            sm.run(eA, eB, gather=False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sm.run(eA, eB)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gates = torch.tensor([float(sm.circ.gates) if sm.circ else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(gates, op=dist.ReduceOp.SUM)
        entry = {"ms": float(t.item()), "n_gpus": world, "levels": sm.circ.levels if sm.circ else 0,
                 "gates": int(gates.item()),
                 "adder": "carry-save tree + prefix addition" if adder == 2 else "ripple (reference schedule)"}
        if rank == 0:
            got = pkg.decrypt_bits(sk, out.cpu().numpy()).reshape(-1, nbits).astype(np.int64)
            C = (got << np.arange(nbits)).sum(-1).reshape(nmat, nmat)
            entry["correct"] = bool(np.array_equal(C, (A @ B) % 2 ** nbits))
        res[name] = entry
        if sm.circ is not None:
            sm.circ.close()
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    pkg = ge.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    p = pkg.default_params()
    n = p.n
    batch = args.batch

    # ---- keys: generated on rank 0, broadcast once over NCCL (NVLink), converted on each GPU
    t0 = time.perf_counter()
    from importlib import import_module

    tdist = import_module("cpu_gpu_tfhe_b200.dist")
    sk = pkg.keygen(2026) if rank == 0 else None
    # NCCL broadcast from rank 0 (no-op at N=1); the LWE key travels too, ONLY so that every rank can
    # decrypt-check its own outputs (explicit opt-in, cpu-gpu-tfhe_b200/dist.py)
    kt = tdist.broadcast_cloud_keys(p, sk, dev, with_secret_key=True)
    d_bk, d_ks, d_key = kt["bk"], kt["ks"], kt["lwe_key"]
    eng = pkg.Engine(device=local)
    eng.load_keys_device(d_bk, d_ks)
    torch.cuda.synchronize()
    key_secs = time.perf_counter() - t0
    lwe_key = d_key.cpu().numpy()
    del d_bk, d_ks

    # ---- synthetic inputs: valid encryptions of random bits under the broadcast key,
    #      generated on the device from a recorded seed
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    alpha = 2.0 ** -15 * (2.0 / np.pi) ** 0.5
    keyt = d_key.to(torch.int64)

    def make_inputs():
        bits = torch.randint(0, 2, (batch,), generator=g, device=dev, dtype=torch.int64)
        a = torch.randint(-2 ** 31, 2 ** 31, (batch, n), generator=g, device=dev, dtype=torch.int64)
        e = torch.round(torch.randn(batch, generator=g, device=dev, dtype=torch.float64) * alpha * 2.0 ** 32).to(torch.int64)
        b = (a * keyt).sum(1) + (2 * bits - 1) * (1 << 29) + e
        s = torch.cat([a, b[:, None]], 1)
        s = ((s + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32).contiguous()
        return bits, s

    bits_a, ca = make_inputs()
    bits_b, cb = make_inputs()
    out = eng.empty(batch)
    expect = 1 - (bits_a & bits_b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps (value) ------------------------------------------------
    for _ in range(args.warmup):
        eng.gate("NAND", ca, cb, out=out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    eng.set_timing(True)
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        eng.gate("NAND", ca, cb, out=out)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    br_ms, ks_ms, calls = eng.get_timing()
    eng.set_timing(False)
    sampler.stop_flag = True
    sampler.join()

    # correctness of the timed output (outside the timed region): decrypt on the device
    ph = (out[:, -1].to(torch.int64) - (out[:, :-1].to(torch.int64) * keyt).sum(1))
    ph = (ph + 2 ** 31) % 2 ** 32 - 2 ** 31
    bits_ok = bool(torch.equal((ph > 0).to(torch.int64), expect))

    # ---- end to end through the C ABI with HOST buffers, copies inside the timed region: all K steps
    #      with page-locked buffers, and one step with ordinary (pageable) memory, which is what a
    #      reference caller's new_gate_bootstrapping_ciphertext_array gives
    h_ca, h_cb = ca.cpu().pin_memory(), cb.cpu().pin_memory()
    h_out = torch.empty((batch, n + 1), dtype=torch.int32).pin_memory()
    np_ca, np_cb, np_out = h_ca.numpy(), h_cb.numpy(), h_out.numpy()
    e2e_steps = args.steps
    eng.gate_host("NAND", np_ca, np_cb, out=np_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.gate_host("NAND", np_ca, np_cb, out=np_out)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    ph = np_out[:, -1].astype(np.int64) - (np_out[:, :-1].astype(np.int64) * lwe_key.astype(np.int64)).sum(1)
    ph = (ph + 2 ** 31) % 2 ** 32 - 2 ** 31
    bits_ok = bits_ok and bool(np.array_equal((ph > 0).astype(np.int64), expect.cpu().numpy()))
    pg_ca, pg_cb, pg_out = np.array(np_ca), np.array(np_cb), np.empty_like(np_out)
    barrier()
    t0 = time.perf_counter()
    eng.gate_host("NAND", pg_ca, pg_cb, out=pg_out)
    torch.cuda.synchronize()
    e2e_pageable_ms = (time.perf_counter() - t0) * 1e3
    bits_ok = bits_ok and bool(np.array_equal(pg_out, np_out))
    del pg_ca, pg_cb, pg_out

    # ---- strong scaling: BASELINE configs[2] read literally, 65536 gates IN TOTAL sharded over the ranks
    lo, hi = tdist.shard_bounds(batch, world, rank)
    mine = hi - lo
    s_ca, s_cb, s_out = ca[:mine], cb[:mine], out[:mine]
    for _ in range(3):
        eng.gate("NAND", s_ca, s_cb, out=s_out)
    barrier()
    sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sv0.record()
    for _ in range(args.steps):
        eng.gate("NAND", s_ca, s_cb, out=s_out)
    sv1.record()
    barrier()
    strong_ms = sv0.elapsed_time(sv1)

    # ---- BASELINE configs[4]: encrypted 16x16 matrix multiply of 8-bit integers over the launched GPUs
    #      (rows of C sharded, operands broadcast once, results gathered; decrypted and checked on rank 0)
    mm = None
    if not args.no_latency:
        mm = matmul_config5(pkg, eng, tdist, sk, p, dev, world, rank, barrier)

    # ---- reduce over ranks: max time ---------------------------------------------------
    tt = torch.tensor([ms, e2e_ms, br_ms, ks_ms, 0.0 if bits_ok else 1.0, strong_ms, e2e_pageable_ms],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_ms, br_ms, ks_ms, bad, strong_ms, e2e_pageable_ms = tt.tolist()

    if rank == 0:
        peak_burst, peak_sust = pkg.measure_fp64_peak(local)
        total_gates = batch * world * args.steps
        value = total_gates / (ms * 1e-3)
        e2e_value = batch * world * e2e_steps / (e2e_ms * 1e-3)
        # roofline of the dominant kernel (blind rotation), per launch: algorithmic flops / event time
        br_ms_per_launch = br_ms / max(1, calls)
        achieved = batch * FLOP_PER_BOOTSTRAP / (br_ms_per_launch * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("blind_rotate_kernel_bytes_per_launch_65536")
            except Exception:
                traffic = None
        latency = circuit_latencies(pkg, eng, sk) if sk is not None and not args.no_latency else None
        if latency is not None and mm is not None:
            latency.update(mm)
        cpu = None
        if not args.no_cpu_baseline and sk is not None and world == 1:
            cores = os.cpu_count() or 1
            probe = cpu_reference_run(sk, pkg, 2)
            gpc = max(4, min(256, int(12.0 / (probe["ms_per_gate_per_core"] / 1e3))))
            cpu = cpu_reference_run(sk, pkg, gpc)
            # single-gate CPU latency = one core's time per bootsNAND (config 1), same run
            cpu["single_gate_ms"] = cpu["ms_per_gate_per_core"]
            # tfhe-spqlios-avx (north_star's CPU path) is not in this image; TFHE's published figure is
            # ~13 ms per gate against ~44 ms for the FFTW build on the same CPU:
            # 3-4x.  The estimate below divides the measured FFTW-shim time by 3.5.
            cpu["spqlios_adjusted"] = {"value": cpu["value"] * 3.5, "unit": UNIT, "factor": 3.5,
                                       "how": "measured reference rate x 3.5 (spqlios-avx vs FFTW path, BASELINE.md section 2); an ESTIMATE"}
            cpu["port_folded"] = cpu_port_folded_rate(sk, pkg, max(4, gpc), cores)
            cpu["circuits"] = cpu_circuit_baseline(sk, pkg, cores)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "%d independent bootstrapped NAND gates per GPU (BASELINE configs[2]), default TFHE gate "
                            "params n=500 N=1024 k=1 l=2 Bgbit=10 ks_t=8 ks_basebit=2" % batch,
                "gates_per_gpu_per_step": batch, "parallelism": "gates sharded over %d GPU(s), no data-path collective; "
                "keys broadcast once over NCCL (%.2f s incl. keygen)" % (world, key_secs),
                "l2": "inputs+outputs 393 MB per step > 126 MB L2 (no flush needed); keys (32.8 MB Fourier BK + 50.3 MB KS) "
                      "are meant to stay L2-resident",
                "bits_decrypt_ok": bool(bad == 0.0),
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * batch * (n + 1) * 4),
                    "d2h_bytes_per_step": int(batch * (n + 1) * 4), "ms_per_step": e2e_ms / e2e_steps,
                    "steps": e2e_steps, "api": "tfhe_b200_gate_host (C ABI, pinned host buffers)",
                    "pageable": {"value": batch * world / (e2e_pageable_ms * 1e-3), "unit": UNIT,
                                 "ms_per_step": e2e_pageable_ms, "steps": 1,
                                 "what": "the same call with ordinary (pageable) host memory"}},
            "strong": {"value": batch * args.steps / (strong_ms * 1e-3), "unit": UNIT,
                       "ms_per_step": strong_ms / args.steps, "gates_total_per_step": batch,
                       "gates_per_gpu_per_step": [h - l for l, h in (tdist.shard_bounds(batch, world, r) for r in range(world))],
                       "what": "BASELINE configs[2] as strong scaling: %d gates in total sharded over %d GPU(s) "
                               "(dist.shard_bounds), device-resident, max over ranks" % (batch, world)},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64", "kernel": "blind_rotate_kernel", "achieved": achieved, "peak": peak_sust,
                "unit": "TFLOP/s", "frac": achieved / peak_sust, "traffic": traffic,
                "peak_burst": peak_burst, "peak_source": "measured live (DFMA micro-benchmark, sustained 0.4 s); "
                "MEASURED_PEAKS.json has no fp64 entry",
                "peak_theoretical": eng.sm_count * 128 * (sampler.result()["sm_max_mhz"] or 1965) * 1e6 / 1e12,
                "frac_of_theoretical": achieved / (eng.sm_count * 128 * (sampler.result()["sm_max_mhz"] or 1965) * 1e6 / 1e12),
                "kernel_ms_per_launch": br_ms_per_launch,
                "flop_per_launch": batch * FLOP_PER_BOOTSTRAP,
                "share_of_step": br_ms / ms, "keyswitch_ms_per_launch": ks_ms / max(1, calls),
            },
            "clocks": sampler.result(),
            "cpu_baseline": cpu,
            "latency": latency,
        }
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-latency", action="store_true", help="skip the single-gate / circuit latency block")
    ap.add_argument("--batch", type=int, default=BATCH, help="gates per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    # banner on rank 0 when NCCL_DEBUG is set in the environment): everything that is not the
    # result line goes to stderr; the line itself is written to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _RESULT_OUT
    _RESULT_OUT = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: at least 3 warm-up steps
        run_b200(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
