"""In-library multi-GPU (tfhe_b200_multi_*, include/tfhe_b200.h): host batches sharded over the
devices of ONE process, keys uploaded once and copied device to device.  The sharding is exercised
on a single GPU too (the same device listed twice = two contexts), the 2-GPU case skips without a
second device.  Also: memory-capped sub-batching of the host-buffer call (boot-gates.cu:2869-2907)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(pkg, sk, count, seed):
    rng = np.random.default_rng(seed)
    ba, bb, bc = (rng.integers(0, 2, count).astype(np.int32) for _ in range(3))
    return ba, bb, bc, pkg.encrypt_bits(sk, ba, seed), pkg.encrypt_bits(sk, bb, seed + 1), pkg.encrypt_bits(sk, bc, seed + 2)


def _check(pkg, sk, devices):
    m = pkg.MultiEngine(devices)
    assert m.ndevices == len(devices)
    m.load_keys(sk.bk, sk.ks)
    count = 1201  # ragged: not a multiple of the device count or of a wave
    ba, bb, bc, ca, cb, cc = _inputs(pkg, sk, count, 40)
    out = m.gate_host("NAND", ca, cb)
    assert np.array_equal(pkg.decrypt_bits(sk, out), 1 - (ba & bb))
    mux = m.mux_host(ca, cb, cc)
    assert np.array_equal(pkg.decrypt_bits(sk, mux), np.where(ba == 1, bb, bc))
    # a gate does not depend on which device (or shard) evaluates it: identical words from one context
    eng = pkg.Engine(device=devices[0])
    eng.load_keys(sk.bk, sk.ks)
    assert np.array_equal(eng.gate_host("NAND", ca, cb), out)
    eng.close()
    assert m.launch_count >= 2 * len(devices)
    # fewer gates than devices: empty shards are fine
    one = m.gate_host("XOR", ca[:1], cb[:1])
    assert pkg.decrypt_bits(sk, one)[0] == ba[0] ^ bb[0]
    m.close()


def test_multi_engine_two_contexts_on_one_gpu(pkg):
    sk = pkg.keygen(31)
    _check(pkg, sk, [0, 0])


def test_multi_engine_two_gpus(pkg):
    if pkg.device_count() < 2:
        pytest.skip("needs two GPUs")
    sk = pkg.keygen(32)
    _check(pkg, sk, [0, 1])


def test_multi_engine_all_devices_default(pkg):
    m = pkg.MultiEngine()
    assert m.ndevices == pkg.device_count()
    m.close()
    with pytest.raises(pkg.EngineError):
        pkg.MultiEngine([pkg.device_count() + 3])


def test_host_call_sub_batches_under_a_memory_cap(pkg):
    """TFHE_B200_HOST_BATCH_LIMIT forces the memory-capped path (a sub-batch of at most one wave
    here): results identical to the uncapped call."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import __graft_entry__ as ge\n"
        "pkg = ge.load_package(); sk = pkg.keygen(33)\n"
        "eng = pkg.Engine(device=0); eng.load_keys(sk.bk, sk.ks)\n"
        "rng = np.random.default_rng(1); n = 1500\n"
        "ba, bb = rng.integers(0, 2, n).astype(np.int32), rng.integers(0, 2, n).astype(np.int32)\n"
        "ca, cb = pkg.encrypt_bits(sk, ba, 1), pkg.encrypt_bits(sk, bb, 2)\n"
        "l0 = eng.launch_count; out = eng.gate_host('AND', ca, cb); l1 = eng.launch_count\n"
        "assert np.array_equal(pkg.decrypt_bits(sk, out), ba & bb)\n"
        "np.save(sys.argv[1], out); print(l1 - l0)\n" % root)
    outs, launches = [], []
    for limit in ("0", "600"):
        env = dict(os.environ, TFHE_B200_HOST_BATCH_LIMIT=limit)
        path = "/tmp/tfhe_b200_cap_%s.npy" % limit
        r = subprocess.run([sys.executable, "-c", code, path], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        launches.append(int(r.stdout.strip().splitlines()[-1]))
        outs.append(np.load(path))
        os.remove(path)
    assert np.array_equal(outs[0], outs[1])
    assert launches[1] > launches[0]  # 1500 gates under a 592-gate cap: three sub-batches
