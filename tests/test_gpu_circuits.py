"""Cipher-level circuits on the GPU: decrypted integers must equal the plaintext results
(bit-exact), for the reference's schedules (BASELINE.json configs 2, 4, 5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def to_bits(vals, nbits):
    vals = np.asarray(vals, dtype=np.int64).reshape(-1)
    return ((vals[:, None] >> np.arange(nbits)) & 1).astype(np.int32)


def from_bits(bits):
    bits = np.asarray(bits, dtype=np.int64)
    return (bits << np.arange(bits.shape[-1])).sum(-1)


def enc_ints(pkg, engine, sk, vals, nbits, seed):
    bits = to_bits(vals, nbits)
    return engine.to_device(pkg.encrypt_bits(sk, bits.reshape(-1), seed))


def dec_ints(pkg, sk, t, nbits):
    return from_bits(pkg.decrypt_bits(sk, t.cpu().numpy()).reshape(-1, nbits))


@pytest.fixture(scope="module")
def sk_engine(pkg):
    sk = pkg.keygen(99)
    eng = pkg.Engine(device=0)
    eng.load_keys(sk.bk, sk.ks)
    yield sk, eng
    eng.close()


def test_carry_operator_gate(pkg, sk_engine):
    """TFHE_B200_GPC = g | (p & c) in one bootstrap, every admissible input (g, p exclusive)."""
    sk, eng = sk_engine
    combos = [(g, p, c) for g in (0, 1) for p in (0, 1) for c in (0, 1) if not (g and p)]
    reps = 40
    g, p, c = [np.repeat(np.array([x[i] for x in combos], np.int32), reps) for i in range(3)]
    out = eng.carry_gate(eng.to_device(pkg.encrypt_bits(sk, g, 21)), eng.to_device(pkg.encrypt_bits(sk, p, 22)),
                         eng.to_device(pkg.encrypt_bits(sk, c, 23)))
    assert np.array_equal(pkg.decrypt_bits(sk, out.cpu().numpy()), g | (p & c))


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_16_bit_addition(pkg, sk_engine, mode):
    """./main 16 a b (BASELINE config 2): both reference schedules, and the prefix adder."""
    sk, eng = sk_engine
    nbits = 16
    a = np.array([12345, 65535, 0, 40000])
    b = np.array([(-6789) & 0xFFFF, 1, 0, 30000])
    circ = pkg.Circuit(eng, "add", nbits, len(a), mode)
    assert circ.levels == {0: 45, 1: 16, 2: 5}[mode]
    out = circ.run(enc_ints(pkg, eng, sk, a, nbits, 1), enc_ints(pkg, eng, sk, b, nbits, 2))
    assert np.array_equal(dec_ints(pkg, sk, out, nbits), (a + b) & 0xFFFF)
    out2 = circ.run(enc_ints(pkg, eng, sk, b, nbits, 3), enc_ints(pkg, eng, sk, a, nbits, 4))  # plans are reusable
    assert np.array_equal(dec_ints(pkg, sk, out2, nbits), (a + b) & 0xFFFF)
    circ.close()


def test_8_bit_multiplication_vector(pkg, sk_engine):
    sk, eng = sk_engine
    nbits = 8
    a = np.array([13, 255, 0, 100, 7])
    b = np.array([11, 255, 99, 3, 36])
    circ = pkg.Circuit(eng, "mul", nbits, len(a))
    out = circ.run(enc_ints(pkg, eng, sk, a, nbits, 5), enc_ints(pkg, eng, sk, b, nbits, 6))
    assert np.array_equal(dec_ints(pkg, sk, out, nbits), (a * b) & 0xFF)
    circ.close()


def test_full_adder_gates(pkg, sk_engine):
    """TFHE_B200_XOR3 / TFHE_B200_MAJ: sum and carry of a full adder, one bootstrap each, all 8 inputs,
    on fresh encryptions AND on bootstrapped outputs (the noise a carry-save tree really feeds them)."""
    sk, eng = sk_engine
    reps = 60
    combos = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    a, b, c = [np.repeat(np.array([x[i] for x in combos], np.int32), reps) for i in range(3)]
    ea, eb, ec = (eng.to_device(pkg.encrypt_bits(sk, v, 70 + i)) for i, v in enumerate((a, b, c)))
    assert np.array_equal(pkg.decrypt_bits(sk, eng.gate3(pkg.binding.XOR3, ea, eb, ec).cpu().numpy()), a ^ b ^ c)
    assert np.array_equal(pkg.decrypt_bits(sk, eng.gate3(pkg.binding.MAJ, ea, eb, ec).cpu().numpy()), (a + b + c) >= 2)
    ba, bb, bc = (eng.gate("AND", t, t) for t in (ea, eb, ec))   # bootstrapped copies: output noise level
    assert np.array_equal(pkg.decrypt_bits(sk, eng.gate3(pkg.binding.XOR3, ba, bb, bc).cpu().numpy()), a ^ b ^ c)
    assert np.array_equal(pkg.decrypt_bits(sk, eng.gate3(pkg.binding.MAJ, ba, bb, bc).cpu().numpy()), (a + b + c) >= 2)


def test_fused_sum_gate(pkg, sk_engine):
    """TFHE_B200_SUMC = a ^ (b | (c & d)) for mutually exclusive b, c (sum bit of the prefix adder fused with
    its last carry operator), one bootstrap, every admissible input, on bootstrapped operands."""
    sk, eng = sk_engine
    combos = [(a, b, c, d) for a in (0, 1) for b in (0, 1) for c in (0, 1) for d in (0, 1) if not (b and c)]
    reps = 40
    v = [np.repeat(np.array([x[i] for x in combos], np.int32), reps) for i in range(4)]
    e = [eng.to_device(pkg.encrypt_bits(sk, v[i], 80 + i)) for i in range(4)]
    e = [eng.gate("AND", t, t) for t in e]   # output-noise level of a previous gate
    out = eng.gate3(pkg.binding.SUMC, e[0], e[1], e[2], d=e[3])
    assert np.array_equal(pkg.decrypt_bits(sk, out.cpu().numpy()), v[0] ^ (v[1] | (v[2] & v[3])))


@pytest.mark.parametrize("adder", [0, 1, 2])
def test_32_bit_multiplication(pkg, sk_engine, adder):
    """BASELINE config 4 (multiplyLweSamples schedule, single precision); adder 1 = prefix tree,
    2 = carry-save tree + one prefix addition."""
    sk, eng = sk_engine
    nbits = 32
    a, b = np.array([40000]), np.array([50000])
    circ = pkg.Circuit(eng, "mul_ex", nbits, 1, adder)
    assert circ.levels == {0: 466, 1: 31, 2: 15}[adder]
    out = circ.run(enc_ints(pkg, eng, sk, a, nbits, 7), enc_ints(pkg, eng, sk, b, nbits, 8))
    assert np.array_equal(dec_ints(pkg, sk, out, nbits), (a * b) & 0xFFFFFFFF)
    circ.close()


def test_carry_save_multiplier_many_operands(pkg, sk_engine):
    """16-bit carry-save products of 24 random pairs incl. the all-ones corner (deep carry chains)."""
    sk, eng = sk_engine
    nbits = 16
    rng = np.random.default_rng(44)
    a, b = rng.integers(0, 2 ** nbits, 24), rng.integers(0, 2 ** nbits, 24)
    a[0] = b[0] = 2 ** nbits - 1
    circ = pkg.Circuit(eng, "mul_ex", nbits, len(a), 2)
    out = circ.run(enc_ints(pkg, eng, sk, a, nbits, 45), enc_ints(pkg, eng, sk, b, nbits, 46))
    assert np.array_equal(dec_ints(pkg, sk, out, nbits), (a * b) & 0xFFFF)
    circ.close()


@pytest.mark.parametrize("adder", [0, 1, 2])
def test_matrix_multiply_4x4_of_8_bit(pkg, sk_engine, adder):
    """Reduced BASELINE config 5 (the reference's own test driver uses 4x4, main.cu:2471)."""
    sk, eng = sk_engine
    nbits, n = 8, 4
    rng = np.random.default_rng(3)
    A = rng.integers(-8, 8, (n, n))
    Bm = rng.integers(-8, 8, (n, n))
    circ = pkg.Circuit(eng, "matmul_ex", n, n, n, nbits, adder)
    out = circ.run(enc_ints(pkg, eng, sk, A.reshape(-1) & 0xFF, nbits, 9),
                   enc_ints(pkg, eng, sk, Bm.reshape(-1) & 0xFF, nbits, 10))
    got = dec_ints(pkg, sk, out, nbits).reshape(n, n)
    assert np.array_equal(got, (A @ Bm) & 0xFF)
    circ.close()


# ---- the rest of the Cipher arithmetic (Cipher.cu:237-630) on real ciphertexts --------------

def sgn(v, nbits):
    v = np.asarray(v, dtype=np.int64) % 2 ** nbits
    return np.where(v >= 2 ** (nbits - 1), v - 2 ** nbits, v)


@pytest.fixture(scope="module")
def operands8():
    a = np.array([100, 3, 255, 0, 128, 127, 77, 200, 13, 250])
    b = np.array([27, 3, 1, 9, 127, 128, 78, 100, 240, 5])
    return a, b


@pytest.mark.parametrize("adder", [0, 1])
def test_subtraction_and_negation(pkg, sk_engine, operands8, adder):
    sk, eng = sk_engine
    a, b = operands8
    ea, eb = enc_ints(pkg, eng, sk, a, 8, 31), enc_ints(pkg, eng, sk, b, 8, 32)
    out = pkg.Circuit(eng, "sub", 8, len(a), adder).run(ea, eb)
    assert np.array_equal(dec_ints(pkg, sk, out, 8), (a - b) & 0xFF)
    if adder == 0:
        out = pkg.Circuit(eng, "neg", 8, len(a)).run(ea)
        assert np.array_equal(dec_ints(pkg, sk, out, 8), (-a) & 0xFF)


@pytest.mark.parametrize("is_signed", [0, 1])
def test_comparisons_min_max(pkg, sk_engine, operands8, is_signed):
    sk, eng = sk_engine
    a, b = operands8
    xa, xb = (sgn(a, 8), sgn(b, 8)) if is_signed else (a, b)
    ea, eb = enc_ints(pkg, eng, sk, a, 8, 33), enc_ints(pkg, eng, sk, b, 8, 34)
    expect = {"GT": xa > xb, "LE": xa <= xb, "LT": xa < xb, "GE": xa >= xb, "EQ": xa == xb, "NE": xa != xb}
    for name, code in pkg.CMP.items():
        out = pkg.Circuit(eng, "compare", 8, len(a), code, is_signed).run(ea, eb)
        assert np.array_equal(pkg.decrypt_bits(sk, out.cpu().numpy()).astype(bool), expect[name]), name
    for want_max in (0, 1):
        out = pkg.Circuit(eng, "minmax", 8, len(a), want_max, is_signed).run(ea, eb)
        assert np.array_equal(dec_ints(pkg, sk, out, 8), np.where((xa > xb) == bool(want_max), a, b))


def test_select_abs_shift(pkg, sk_engine, operands8):
    sk, eng = sk_engine
    a, b = operands8
    ea, eb = enc_ints(pkg, eng, sk, a, 8, 35), enc_ints(pkg, eng, sk, b, 8, 36)
    sel = (np.arange(len(a)) % 2).astype(np.int32)
    es = eng.to_device(pkg.encrypt_bits(sk, sel, 37))
    out = pkg.Circuit(eng, "select", 8, len(a)).run(es, ea, eb)
    assert np.array_equal(dec_ints(pkg, sk, out, 8), np.where(sel == 1, a, b))
    out = pkg.Circuit(eng, "abs", 8, len(a), 1).run(ea)
    assert np.array_equal(dec_ints(pkg, sk, out, 8), np.abs(sgn(a, 8)) & 0xFF)
    out = pkg.Circuit(eng, "shift", 8, len(a), 3, pkg.SHIFT["RIGHT_ARITH"]).run(ea)
    assert np.array_equal(dec_ints(pkg, sk, out, 8), (sgn(a, 8) >> 3) & 0xFF)
    out = pkg.Circuit(eng, "shift", 8, len(a), 2, pkg.SHIFT["LEFT"]).run(ea)
    assert np.array_equal(dec_ints(pkg, sk, out, 8), (a << 2) & 0xFF)


@pytest.mark.parametrize("is_signed", [0, 1])
def test_division(pkg, sk_engine, operands8, is_signed):
    """operator/ (Cipher.cu:494): restoring division, quotient and remainder."""
    sk, eng = sk_engine
    a, b = operands8
    ea, eb = enc_ints(pkg, eng, sk, a, 8, 38), enc_ints(pkg, eng, sk, b, 8, 39)
    out = dec_ints(pkg, sk, pkg.Circuit(eng, "div", 8, len(a), is_signed, 1).run(ea, eb), 8).reshape(-1, 2)
    if is_signed:
        xa, xb = sgn(a, 8), sgn(b, 8)
        q, r = np.sign(xa) * np.sign(xb) * (np.abs(xa) // np.abs(xb)), np.abs(xa) % np.abs(xb)
    else:
        q, r = a // b, a % b
    assert np.array_equal(out[:, 0], q & 0xFF) and np.array_equal(out[:, 1], r & 0xFF)


def test_karatsuba_and_full_products(pkg, sk_engine):
    """karatMasterSuba (main.cu:1866) and the double-precision product on ciphertexts: full 16-bit
    products of 8-bit operands."""
    sk, eng = sk_engine
    a = np.array([255, 200, 17, 0, 128])
    b = np.array([255, 3, 250, 99, 128])
    ea, eb = enc_ints(pkg, eng, sk, a, 8, 41), enc_ints(pkg, eng, sk, b, 8, 42)
    for kind in ("mul_karatsuba", "mul_full"):
        out = pkg.Circuit(eng, kind, 8, len(a), 1).run(ea, eb)
        assert np.array_equal(dec_ints(pkg, sk, out, 16), a * b), kind


def test_cannon_matrix_multiply_3x3(pkg, sk_engine):
    """BOOTS_CannonsAlgo (main.cu:2590) as a plan."""
    sk, eng = sk_engine
    nbits, n = 8, 3
    rng = np.random.default_rng(11)
    A, Bm = rng.integers(-8, 8, (n, n)), rng.integers(-8, 8, (n, n))
    circ = pkg.Circuit(eng, "matmul_cannon", n, nbits, 1)
    out = circ.run(enc_ints(pkg, eng, sk, A.reshape(-1) & 0xFF, nbits, 43),
                   enc_ints(pkg, eng, sk, Bm.reshape(-1) & 0xFF, nbits, 44))
    assert np.array_equal(dec_ints(pkg, sk, out, nbits).reshape(n, n), (A @ Bm) & 0xFF)


def test_plan_replays_a_cuda_graph_with_identical_results(pkg, sk_engine):
    """The launch sequence of a plan is captured into a CUDA graph at the first run on a non-default
    stream and replayed afterwards: same ciphertext words as direct launches (the kernels are
    deterministic), launch accounting intact."""
    import torch

    sk, eng = sk_engine
    nbits = 16
    a, b = np.array([12345, 7]), np.array([(-6789) & 0xFFFF, 65530])
    ea, eb = enc_ints(pkg, eng, sk, a, nbits, 51), enc_ints(pkg, eng, sk, b, nbits, 52)
    circ = pkg.Circuit(eng, "add", nbits, len(a), 2)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        circ.set_graph(False)
        l0 = eng.launch_count
        direct = circ.run(ea, eb).clone()
        per_run = eng.launch_count - l0
        assert not circ.used_graph
        circ.set_graph(True)
        first = circ.run(ea, eb).clone()     # captures, then replays
        assert circ.used_graph
        l1 = eng.launch_count
        again = circ.run(ea, eb).clone()
        assert circ.used_graph and eng.launch_count - l1 == per_run
    st.synchronize()
    assert torch.equal(direct, first) and torch.equal(direct, again)
    assert np.array_equal(dec_ints(pkg, sk, again, nbits), (a + b) & 0xFFFF)
    # on the legacy default stream (not capturable) the plan launches directly
    out = circ.run(ea, eb)
    assert not circ.used_graph and torch.equal(out, direct)


def test_independent_plans_merged_level_by_level(pkg, sk_engine):
    """tfhe_b200_circuit_run_many: a 16-bit prefix adder, an 8-bit ripple adder, a comparison and a
    multiplier run TOGETHER, their levels sharing launches; same words as running them one by one,
    with far fewer launches."""
    import torch

    sk, eng = sk_engine
    specs = [("add", (16, 2, 2), [np.array([1234, 65535]), np.array([4321, 1])], 16, lambda x, y: (x + y) & 0xFFFF),
             ("add", (8, 3, 0), [np.array([200, 17, 0]), np.array([100, 3, 255])], 8, lambda x, y: (x + y) & 0xFF),
             ("sub", (12, 1, 1), [np.array([100]), np.array([3000])], 12, lambda x, y: (x - y) & 0xFFF),
             ("mul_ex", (6, 2, 1), [np.array([13, 63]), np.array([5, 63])], 6, lambda x, y: (x * y) & 0x3F)]
    plans, operands = [], []
    for i, (kind, args, vals, nbits, _) in enumerate(specs):
        plans.append(pkg.Circuit(eng, kind, *args))
        operands.append([enc_ints(pkg, eng, sk, v, nbits, 60 + 2 * i + j) for j, v in enumerate(vals)])
    for p in plans:
        p.set_graph(False)
    l0 = eng.launch_count
    single = [p.run(*ops).clone() for p, ops in zip(plans, operands)]
    l1 = eng.launch_count
    merged = pkg.Circuit.run_many(plans, operands)
    l2 = eng.launch_count
    torch.cuda.synchronize()
    for (kind, args, vals, nbits, f), s, m in zip(specs, single, merged):
        assert torch.equal(s, m), kind
        assert np.array_equal(dec_ints(pkg, sk, m, nbits), f(*vals)), kind
    assert (l2 - l1) < 0.6 * (l1 - l0), (l1 - l0, l2 - l1)
    # the same plan twice in one merged run is refused (one workspace per plan)
    with pytest.raises(pkg.EngineError):
        pkg.Circuit.run_many([plans[0], plans[0]], [operands[0], operands[0]])
