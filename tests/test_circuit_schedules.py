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
                  # prefix adder: g/p level + scan levels, the last of which produces the sum bits (SUMC)
                  2: (1 + int(np.ceil(np.log2(nbits - 1)))) if nbits > 2 else 2}[mode]
        assert circ.levels == expect
    circ.close()


def test_16_bit_depths(pkg):
    """BASELINE config 2: the reference's schedules are 45 / 16 sequential bootstrap batches,
    the prefix adder 5 (one level of generate / propagate, four scan levels of which the last one
    produces the sum bits)."""
    assert [pkg.Circuit(None, "add", 16, 1, m).levels for m in (0, 1, 2)] == [45, 16, 5]


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
    """BASELINE config 4: 1 + 5 adder-tree levels; ripple 5*93 = 465 levels, prefix 5*6 = 30."""
    ripple = pkg.Circuit(None, "mul_ex", 32, 1, 0)
    prefix = pkg.Circuit(None, "mul_ex", 32, 1, 1)
    assert ripple.levels == 1 + 5 * 93 and prefix.levels == 1 + 5 * 6
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


# ---- the rest of the Cipher arithmetic (Cipher.cu:237-630) -------------------------------------

def signed(v, nbits):
    v = np.asarray(v, dtype=np.int64) % 2 ** nbits
    return np.where(v >= 2 ** (nbits - 1), v - 2 ** nbits, v)


def all_pairs(nbits):
    v = np.arange(2 ** nbits)
    a, b = np.meshgrid(v, v, indexing="ij")
    return a.reshape(-1), b.reshape(-1)


@pytest.mark.parametrize("adder", [0, 1])
@pytest.mark.parametrize("nbits", [1, 2, 4, 5])
def test_subtraction_exhaustive(pkg, nbits, adder):
    a, b = all_pairs(nbits)
    circ = pkg.Circuit(None, "sub", nbits, len(a), adder)
    out = circ.simulate(to_bits(a, nbits), to_bits(b, nbits))
    assert np.array_equal(from_bits(out, nbits), (a - b) % 2 ** nbits)


@pytest.mark.parametrize("nbits", [1, 2, 3, 6, 9])
def test_negation_exhaustive(pkg, nbits):
    a = np.arange(2 ** nbits)
    circ = pkg.Circuit(None, "neg", nbits, len(a))
    assert np.array_equal(from_bits(circ.simulate(to_bits(a, nbits)), nbits), (-a) % 2 ** nbits)
    if nbits == 9:
        assert circ.levels == 1 + 3  # OR scan over 8 positions + the XOR level


@pytest.mark.parametrize("is_signed", [0, 1])
@pytest.mark.parametrize("nbits", [1, 2, 3, 5])
def test_comparisons_exhaustive(pkg, nbits, is_signed):
    a, b = all_pairs(nbits)
    xa, xb = (signed(a, nbits), signed(b, nbits)) if is_signed else (a, b)
    expect = {"GT": xa > xb, "LE": xa <= xb, "LT": xa < xb, "GE": xa >= xb, "EQ": xa == xb, "NE": xa != xb}
    for name, code in pkg.CMP.items():
        circ = pkg.Circuit(None, "compare", nbits, len(a), code, is_signed)
        out = circ.simulate(to_bits(a, nbits), to_bits(b, nbits))
        assert np.array_equal(out.astype(bool), expect[name]), name


def test_comparison_depth(pkg):
    """16-bit a > b: 1 + log2(16) bootstrap levels (the reference chains 16 x 4 gates, Cipher.cu:561-598)."""
    assert pkg.Circuit(None, "compare", 16, 1, pkg.CMP["GT"], 1).levels == 5
    assert pkg.Circuit(None, "compare", 16, 1, pkg.CMP["EQ"], 0).levels == 5


@pytest.mark.parametrize("is_signed", [0, 1])
@pytest.mark.parametrize("want_max", [0, 1])
def test_min_max_exhaustive(pkg, want_max, is_signed):
    nbits = 4
    a, b = all_pairs(nbits)
    xa, xb = (signed(a, nbits), signed(b, nbits)) if is_signed else (a, b)
    circ = pkg.Circuit(None, "minmax", nbits, len(a), want_max, is_signed)
    out = from_bits(circ.simulate(to_bits(a, nbits), to_bits(b, nbits)), nbits)
    expect = np.where((xa > xb) == bool(want_max), a, b)
    assert np.array_equal(out, expect)


def test_select(pkg):
    nbits, count = 6, 50
    rng = np.random.default_rng(5)
    sel = rng.integers(0, 2, count)
    a, b = rng.integers(0, 64, count), rng.integers(0, 64, count)
    circ = pkg.Circuit(None, "select", nbits, count)
    out = from_bits(circ.simulate(sel, to_bits(a, nbits), to_bits(b, nbits)), nbits)
    assert np.array_equal(out, np.where(sel == 1, a, b))
    assert circ.levels == 1


@pytest.mark.parametrize("adder", [0, 1])
def test_absolute_exhaustive(pkg, adder):
    nbits = 6
    a = np.arange(2 ** nbits)
    circ = pkg.Circuit(None, "abs", nbits, len(a), adder)
    out = from_bits(circ.simulate(to_bits(a, nbits)), nbits)
    assert np.array_equal(out, np.abs(signed(a, nbits)) % 2 ** nbits)


@pytest.mark.parametrize("kind", ["LEFT", "RIGHT_LOGICAL", "RIGHT_ARITH"])
@pytest.mark.parametrize("amount", [0, 1, 3, 8, 11])
def test_shifts(pkg, kind, amount):
    nbits = 8
    a = np.arange(2 ** nbits)
    circ = pkg.Circuit(None, "shift", nbits, len(a), amount, pkg.SHIFT[kind])
    out = from_bits(circ.simulate(to_bits(a, nbits)), nbits)
    if kind == "LEFT":
        expect = (a << amount) % 2 ** nbits
    elif kind == "RIGHT_LOGICAL":
        expect = a >> amount
    else:
        expect = (signed(a, nbits) >> min(amount, 63)) % 2 ** nbits
    assert np.array_equal(out, expect)
    assert circ.levels == 0 and circ.gates == 0  # bootstrap free


@pytest.mark.parametrize("adder", [0, 1])
@pytest.mark.parametrize("is_signed", [0, 1])
def test_division_exhaustive(pkg, is_signed, adder):
    nbits = 4
    a, b = all_pairs(nbits)
    keep = b != 0
    a, b = a[keep], b[keep]
    circ = pkg.Circuit(None, "div", nbits, len(a), is_signed, adder)
    out = from_bits(circ.simulate(to_bits(a, nbits), to_bits(b, nbits)), nbits).reshape(-1, 2)
    if is_signed:
        xa, xb = signed(a, nbits), signed(b, nbits)
        q = np.sign(xa) * np.sign(xb) * (np.abs(xa) // np.abs(xb))
        r = np.abs(xa) % np.abs(xb)
    else:
        q, r = a // b, a % b
    assert np.array_equal(out[:, 0], q % 2 ** nbits)
    assert np.array_equal(out[:, 1], r % 2 ** nbits)


def test_division_16_bit_random(pkg):
    nbits, count = 16, 40
    rng = np.random.default_rng(8)
    a = rng.integers(0, 2 ** nbits, count)
    b = rng.integers(1, 2 ** nbits, count)
    b[:10] = rng.integers(1, 40, 10)
    circ = pkg.Circuit(None, "div", nbits, count, 0, 1)
    out = from_bits(circ.simulate(to_bits(a, nbits), to_bits(b, nbits)), nbits).reshape(-1, 2)
    assert np.array_equal(out[:, 0], a // b) and np.array_equal(out[:, 1], a % b)


@pytest.mark.parametrize("adder", [0, 1])
@pytest.mark.parametrize("kind", ["mul_full", "mul_karatsuba"])
def test_full_precision_products_exhaustive(pkg, kind, adder):
    """Double-precision product (BOOTS_vectorMultiplication isDoublePrecision) and one level of
    Karatsuba (karatMasterSuba main.cu:1866): all pairs of 4-bit operands, and random 8/16-bit ones."""
    a, b = all_pairs(4)
    circ = pkg.Circuit(None, kind, 4, len(a), adder)
    out = from_bits(circ.simulate(to_bits(a, 4), to_bits(b, 4)), 8)
    assert np.array_equal(out, a * b)
    rng = np.random.default_rng(4)
    for nbits in (8, 16):
        a, b = rng.integers(0, 2 ** nbits, 20), rng.integers(0, 2 ** nbits, 20)
        a[0] = b[0] = 2 ** nbits - 1
        circ = pkg.Circuit(None, kind, nbits, len(a), adder)
        out = from_bits(circ.simulate(to_bits(a, nbits), to_bits(b, nbits)), 2 * nbits)
        assert np.array_equal(out, a * b), nbits


def test_karatsuba_uses_fewer_gates_at_32_bits(pkg):
    full = pkg.Circuit(None, "mul_full", 32, 1, 1)
    kar = pkg.Circuit(None, "mul_karatsuba", 32, 1, 1)
    assert kar.gates < full.gates


@pytest.mark.parametrize("adder", [0, 1])
def test_cannon_matrix_multiply(pkg, adder):
    nbits = 8
    for n in (1, 2, 3, 4):
        rng = np.random.default_rng(n)
        A, B = rng.integers(-8, 8, (n, n)), rng.integers(-8, 8, (n, n))
        circ = pkg.Circuit(None, "matmul_cannon", n, nbits, adder)
        out = circ.simulate(to_bits(A.reshape(-1) & 0xFF, nbits), to_bits(B.reshape(-1) & 0xFF, nbits))
        assert np.array_equal(from_bits(out, nbits).reshape(n, n), (A @ B) & 0xFF), n


@pytest.mark.parametrize("nbits", [2, 3, 4, 5, 8, 16, 32])
def test_carry_save_multiplication_schedule(pkg, nbits):
    """TFHE_B200_ADDER_CARRY_SAVE: partial products -> Wallace tree of full adders (XOR3 / MAJ, one level
    per 3:2 compression) -> one prefix addition.  Exhaustive for small widths, random + corner cases else."""
    if nbits <= 4:
        a, b = np.meshgrid(np.arange(2 ** nbits), np.arange(2 ** nbits))
        a, b = a.reshape(-1), b.reshape(-1)
    else:
        rng = np.random.default_rng(nbits)
        a = rng.integers(0, 2 ** nbits, 40)
        b = rng.integers(0, 2 ** nbits, 40)
        a[0] = b[0] = 2 ** nbits - 1
        a[1], b[1] = 2 ** nbits - 1, 1
        a[2], b[2] = 0, 2 ** nbits - 1
    circ = pkg.Circuit(None, "mul_ex", nbits, len(a), 2)
    out = circ.simulate(to_bits(a, nbits), to_bits(b, nbits))
    assert np.array_equal(from_bits(out, nbits), (a.astype(object) * b.astype(object)) % 2 ** nbits)
    circ.close()


def test_carry_save_depth_and_size(pkg):
    """32 bits: 1 AND level + 8 compression levels + 6 levels of the final prefix addition = 15, with a
    quarter of the gates of the prefix-adder tree (BASELINE config 4)."""
    cs, prefix = pkg.Circuit(None, "mul_ex", 32, 1, 2), pkg.Circuit(None, "mul_ex", 32, 1, 1)
    assert cs.levels == 15 and prefix.levels == 31
    assert cs.gates < prefix.gates / 3
    # per-level widths (tfhe_b200_circuit_level_gates): the AND level, then the narrow reduction levels
    for c in (cs, prefix):
        assert len(c.level_gates) == c.levels and sum(c.level_gates) == c.gates
    assert cs.level_gates[0] == 32 * 33 // 2 and max(cs.level_gates[1:]) < 200
    a, b = np.array([40000]), np.array([50000])
    assert from_bits(cs.simulate(to_bits(a, 32), to_bits(b, 32)), 32)[0] == (40000 * 50000) % 2 ** 32


def test_carry_save_matrix_multiply_schedule(pkg):
    nbits, n = 8, 5
    rng = np.random.default_rng(13)
    A = rng.integers(-8, 8, (n, n + 1))
    B = rng.integers(-8, 8, (n + 1, n - 1))
    circ = pkg.Circuit(None, "matmul_ex", n, n + 1, n - 1, nbits, 2)
    out = circ.simulate(to_bits(A.reshape(-1) & 0xFF, nbits), to_bits(B.reshape(-1) & 0xFF, nbits))
    assert np.array_equal(from_bits(out, nbits).reshape(n, n - 1), (A @ B) & 0xFF)
    ripple = pkg.Circuit(None, "matmul_ex", n, n + 1, n - 1, nbits, 0)
    assert circ.gates < ripple.gates / 2 and circ.levels < ripple.levels / 4
    big = pkg.Circuit(None, "matmul_ex", 16, 16, 16, 8, 2)   # BASELINE config 5
    print("16x16x16 8-bit matmul, carry-save: %d levels, %d gates" % (big.levels, big.gates))
