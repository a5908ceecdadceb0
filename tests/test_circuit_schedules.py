"""Gate schedules of the Cipher-level circuits, checked on PLAINTEXT bits on the host
(tfhe_b200_circuit_simulate): every level structure the GPU runs (test_gpu_circuits.py) is
first proven to compute the right integers, for the reference's schedules
(taskLevelParallelAdd[_bitwise] main.cu:619/821, multiplyLweSamples :1483,
BOOTS_matrixMultiplication :2342) and for the depth-reduced parallel-prefix variants."""
import numpy as np
import pytest


def to_bits(vals, nbits):
    vals = np.asarray(vals, dtype=np.int64).reshape(-1)
    return ((vals[:, None] >> np.arange(nbits)) & 1).astype(np.int32)


def from_bits(bits, nbits):
    bits = np.asarray(bits, dtype=np.int64).reshape(-1, nbits)
    return (bits << np.arange(nbits)).sum(-1)


@pytest.mark.parametrize("nbits", [1, 2, 3, 5, 8, 16, 32])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_addition_schedules(pkg, nbits, mode):
    rng = np.random.default_rng(100 * nbits + mode)
    count = 37
    a = rng.integers(0, 2 ** nbits, count)
    b = rng.integers(0, 2 ** nbits, count)
    a[0], b[0] = 2 ** nbits - 1, 1            # full carry chain
    a[1], b[1] = 2 ** nbits - 1, 2 ** nbits - 1
    a[2], b[2] = 0, 0
    circ = pkg.Circuit(None, "add", nbits, count, mode)
    out = circ.simulate(to_bits(a, nbits), to_bits(b, nbits))
    assert np.array_equal(from_bits(out, nbits), (a + b) % 2 ** nbits)
    if nbits > 1:
        expect = {0: 3 * nbits - 3, 1: nbits,
                  2: 2 + (int(np.ceil(np.log2(nbits - 1))) if nbits > 2 else 0)}[mode]
        assert circ.levels == expect
    circ.close()


def test_16_bit_depths(pkg):
    """BASELINE config 2: the reference's schedules are 45 / 16 sequential bootstrap batches,
    the prefix adder 6."""
    assert [pkg.Circuit(None, "add", 16, 1, m).levels for m in (0, 1, 2)] == [45, 16, 6]


@pytest.mark.parametrize("adder", [0, 1])
@pytest.mark.parametrize("nbits", [2, 4, 8, 16])
def test_multiplication_schedules(pkg, nbits, adder):
    rng = np.random.default_rng(nbits + adder)
    count = 9
    a = rng.integers(0, 2 ** nbits, count)
    b = rng.integers(0, 2 ** nbits, count)
    a[0] = b[0] = 2 ** nbits - 1
    circ = pkg.Circuit(None, "mul_ex", nbits, count, adder)
    out = circ.simulate(to_bits(a, nbits), to_bits(b, nbits))
    assert np.array_equal(from_bits(out, nbits), (a * b) % 2 ** nbits)
    circ.close()


def test_32_bit_multiplication_depth(pkg):
    """BASELINE config 4: 1 + 5 adder-tree levels; ripple 5*93 = 465 levels, prefix 5*7 = 35."""
    ripple = pkg.Circuit(None, "mul_ex", 32, 1, 0)
    prefix = pkg.Circuit(None, "mul_ex", 32, 1, 1)
    assert ripple.levels == 1 + 5 * 93 and prefix.levels == 1 + 5 * 7
    a, b = np.array([40000]), np.array([50000])
    for c in (ripple, prefix):
        assert from_bits(c.simulate(to_bits(a, 32), to_bits(b, 32)), 32)[0] == (40000 * 50000) % 2 ** 32


@pytest.mark.parametrize("adder", [0, 1])
def test_matrix_multiply_schedule(pkg, adder):
    nbits, n = 8, 5
    rng = np.random.default_rng(3 + adder)
    A = rng.integers(-8, 8, (n, n + 1))
    B = rng.integers(-8, 8, (n + 1, n - 1))
    circ = pkg.Circuit(None, "matmul_ex", n, n + 1, n - 1, nbits, adder)
    out = circ.simulate(to_bits(A.reshape(-1) & 0xFF, nbits), to_bits(B.reshape(-1) & 0xFF, nbits))
    assert np.array_equal(from_bits(out, nbits).reshape(n, n - 1), (A @ B) & 0xFF)
    circ.close()


def test_plan_without_engine_cannot_run(pkg):
    circ = pkg.Circuit(None, "add", 4, 1, 2)
    assert circ.eng is None and circ.gates > 0
