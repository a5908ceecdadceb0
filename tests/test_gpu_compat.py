"""Drop-in boundary test: the REFERENCE'S OWN objects (key sets, LweSample, TLweSample,
TorusPolynomial built by the reference's code in oracle/_ref) are handed, unchanged, to the
same-named entry points of libtfhe_b200.so (include/tfhe_compat.h).  Proves the struct
layouts and the calling conventions match; results are checked against the reference's own
functions on the same objects.  Needs oracle/_ref (travels with the repo) and a GPU."""
import ctypes

import numpy as np
import pytest

from conftest import wrap32
from oracle.pyoracle import GATES, Oracle, Ref, have_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")]

vp = ctypes.c_void_p
MU = 0x20000000


@pytest.fixture(scope="module")
def world(pkg):
    o = Oracle()
    r = Ref().keygen((314, 1592, 657))
    L = ctypes.CDLL(pkg.lib_path())  # RTLD_LOCAL: its bootsNAND does not clash with the reference's
    R = r.L
    for f in ("ref_cloud_keyset", "ref_bkfft", "ref_tgsw_fft_array", "ref_tgsw_params", "ref_ks_key",
              "ref_sample_new", "ref_tlwe_new", "ref_torus_poly_new"):
        getattr(R, f).restype = vp
    L.tfhe_b200_keys_to_gpu.restype = vp
    L.convertBitToNumberZero_GPU.restype = vp
    keys = r.export_keys(o.params)
    return o, r, R, L, keys


class Sample:
    """A reference LweSample object (allocated and freed by the reference)."""

    def __init__(self, R, h, n, flat=None, extracted=0):
        self.R, self.n = R, n
        self.p = vp(R.ref_sample_new(h, extracted))
        if flat is not None:
            flat = np.ascontiguousarray(flat, np.int32)
            R.ref_sample_set(self.p, flat.ctypes.data_as(vp), n)

    def get(self):
        out = np.zeros(self.n + 1, np.int32)
        self.R.ref_sample_get(self.p, out.ctypes.data_as(vp), self.n)
        return out

    def __del__(self):
        self.R.ref_sample_free(self.p)


def test_classic_gates_on_reference_objects(world):
    o, r, R, L, keys = world
    cloud = vp(R.ref_cloud_keyset(r.h))
    bits = [(0, 0), (0, 1), (1, 0), (1, 1)]
    truth = {"NAND": lambda a, b: 1 - (a & b), "OR": lambda a, b: a | b, "AND": lambda a, b: a & b,
             "XOR": lambda a, b: a ^ b, "XNOR": lambda a, b: 1 - (a ^ b), "NOR": lambda a, b: 1 - (a | b),
             "ANDNY": lambda a, b: (1 - a) & b, "ANDYN": lambda a, b: a & (1 - b),
             "ORNY": lambda a, b: (1 - a) | b, "ORYN": lambda a, b: a | (1 - b)}
    for g in GATES:
        for a, b in bits:
            ca, cb = r.encrypt(a), r.encrypt(b)
            sa, sb, res = Sample(R, r.h, 500, ca), Sample(R, r.h, 500, cb), Sample(R, r.h, 500)
            getattr(L, "boots" + g)(res.p, sa.p, sb.p, cloud)
            got = res.get()
            ref = r.gate(g, ca, cb)  # the reference's own bootsXXX on the same inputs
            assert int(r.phase(got) > 0) == truth[g](a, b) == int(r.phase(ref) > 0), (g, a, b)
            assert abs(int(wrap32(r.phase(got) - r.phase(ref)))) / 2.0 ** 32 < 2.0 ** -5
    # in-place: result aliases the first operand (Cipher.cu:373)
    sa, sb = Sample(R, r.h, 500, r.encrypt(1)), Sample(R, r.h, 500, r.encrypt(0))
    L.bootsAND(sa.p, sa.p, sb.p, cloud)
    assert r.phase(sa.get()) < 0
    # MUX, NOT, COPY, CONSTANT
    for a in (0, 1):
        sa, sb, sc, res = (Sample(R, r.h, 500, r.encrypt(x)) for x in (a, 1, 0, 0))
        L.bootsMUX(res.p, sa.p, sb.p, sc.p, cloud)
        assert int(r.phase(res.get()) > 0) == (1 if a else 0)
    s1, res = Sample(R, r.h, 500, r.encrypt(1)), Sample(R, r.h, 500)
    L.bootsNOT(res.p, s1.p, cloud)
    assert np.array_equal(res.get(), (-s1.get().astype(np.int64)).astype(np.int32))
    L.bootsCOPY(res.p, s1.p, cloud)
    assert np.array_equal(res.get(), s1.get())
    L.bootsCONSTANT(res.p, 1, cloud)
    assert res.get()[-1] == MU and not res.get()[:-1].any()


def test_bootstrap_and_keyswitch_on_reference_objects(world):
    o, r, R, L, keys = world
    bkfft = vp(R.ref_bkfft(r.h))
    x = o.gate_prologue("NAND", r.encrypt(1), r.encrypt(1))
    sx, u = Sample(R, r.h, 500, x), Sample(R, r.h, 1024, extracted=1)
    L.tfhe_bootstrap_woKS_FFT(u.p, bkfft, ctypes.c_int32(MU), sx.p)
    ref_u = r.bootstrap_woks(MU, x)
    ph_g, ph_r = o.phase(keys.tlwe_key, u.get()), o.phase(keys.tlwe_key, ref_u)
    assert ph_g < 0 and ph_r < 0 and abs(int(wrap32(ph_g - ph_r))) / 2.0 ** 32 < 2.0 ** -5
    # key switch of the reference's own u: integer arithmetic, must be bit-exact
    su, res = Sample(R, r.h, 1024, ref_u, extracted=1), Sample(R, r.h, 500)
    L.lweKeySwitch(res.p, vp(R.ref_ks_key(r.h)), su.p)
    assert np.array_equal(res.get(), r.keyswitch(ref_u))
    # bootstrap with key switch
    L.tfhe_bootstrap_FFT(res.p, bkfft, ctypes.c_int32(MU), sx.p)
    assert r.phase(res.get()) < 0


def test_extern_mul_and_blind_rotate_on_reference_objects(world):
    o, r, R, L, keys = world
    rng = np.random.default_rng(8)
    acc = rng.integers(-2 ** 31, 2 ** 31, (2, 1024), dtype=np.int64).astype(np.int32)
    params = vp(R.ref_tgsw_params(r.h))
    base = R.ref_tgsw_fft_array(r.h)
    # sizeof(TGswSampleFFT) = 2 pointers + 2 ints = 24 bytes (tgsw.h:78-84)
    idx = 7
    t = vp(R.ref_tlwe_new(r.h))
    R.ref_tlwe_set(t, acc.ctypes.data_as(vp), 1024, 1)
    L.tGswFFTExternMulToTLwe(t, vp(base + 24 * idx), params)
    got = np.zeros_like(acc)
    R.ref_tlwe_get(t, got.ctypes.data_as(vp), 1024, 1)
    exact = o.extern_mul_exact(keys.bk[idx], acc)
    assert np.abs(wrap32(got.astype(np.int64) - exact.astype(np.int64))).max() <= 1
    assert np.abs(wrap32(got.astype(np.int64) - r.extern_mul(idx, acc).astype(np.int64))).max() <= 2
    # tfhe_blindRotateAndExtract_FFT with n = 0 iterations is pure integer work: bit-exact
    tv = rng.integers(-2 ** 31, 2 ** 31, 1024, dtype=np.int64).astype(np.int32)
    pv = vp(R.ref_torus_poly_new(1024, tv.ctypes.data_as(vp)))
    u = Sample(R, r.h, 1024, extracted=1)
    bara = (ctypes.c_int * 1)(0)
    L.tfhe_blindRotateAndExtract_FFT(u.p, pv, vp(base), 1500, bara, 0, params)
    expect = o.ctx(keys).blind_rotate_and_extract(tv, 1500, np.zeros(0, np.int32))
    assert np.array_equal(u.get(), expect)
    # a short blind rotation of a reference accumulator
    acc0 = np.zeros((2, 1024), np.int32)
    acc0[1] = MU
    R.ref_tlwe_set(t, acc0.ctypes.data_as(vp), 1024, 1)
    bara = np.array([3, 2047, 0, 1024, 77], np.int32)
    L.tfhe_blindRotate_FFT(t, vp(base), bara.ctypes.data_as(vp), 5, params)
    R.ref_tlwe_get(t, got.ctypes.data_as(vp), 1024, 1)
    ref_acc = r.blind_rotate(acc0, bara)

    def phase(a):
        full = np.convolve(a[0].astype(np.int64), keys.tlwe_key.astype(np.int64))
        rr = full[:1024].copy()
        rr[:1023] -= full[1024:]
        return wrap32(a[1].astype(np.int64) - rr)

    assert np.abs(wrap32(phase(got) - phase(ref_acc))).max() / 2.0 ** 32 < 2.0 ** -8
    R.ref_tlwe_free(t)
    R.ref_torus_poly_free(pv)


def test_batched_fullgpu_family(world):
    """LweSample_16 convention: a on the device, b on the host (boot-gates.cu:462-476)."""
    import torch

    o, r, R, L, keys = world

    class S16(ctypes.Structure):
        _fields_ = [("a", vp), ("b", ctypes.POINTER(ctypes.c_int)), ("cv", ctypes.POINTER(ctypes.c_double))]

    cloud = vp(R.ref_cloud_keyset(r.h))
    handle = vp(L.tfhe_b200_keys_to_gpu(cloud))
    nb = 6
    rng = np.random.default_rng(5)
    bits = [rng.integers(0, 2, nb) for _ in range(4)]

    torch.zeros(1, device="cuda")
    cudart = ctypes.CDLL("libcudart.so.12")  # the runtime torch has already loaded
    cudart.cudaMemcpy.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_int]

    def make(bitvec):
        s = ctypes.cast(L.convertBitToNumberZero_GPU(nb, cloud), ctypes.POINTER(S16))
        flat = np.stack([r.encrypt(int(b)) for b in bitvec])
        a = np.ascontiguousarray(flat[:, :-1])
        assert cudart.cudaMemcpy(s.contents.a, a.ctypes.data_as(vp), a.nbytes, 1) == 0
        for i in range(nb):
            s.contents.b[i] = int(flat[i, -1])
        return s, flat

    def read(s, count):
        a = np.zeros((count, 500), np.int32)
        assert cudart.cudaMemcpy(a.ctypes.data_as(vp), s.contents.a, a.nbytes, 2) == 0
        b = np.array([s.contents.b[i] for i in range(count)], np.int32)
        return np.concatenate([a, b[:, None]], 1)

    sa, fa = make(bits[0])
    sb, fb = make(bits[1])
    sc, fc = make(bits[2])
    res = ctypes.cast(L.convertBitToNumberZero_GPU(2 * nb, cloud), ctypes.POINTER(S16))
    L.bootsAND_fullGPU_n_Bit(res, sa, sb, nb, handle, None, None)
    dec = (o.phases(keys.lwe_key, read(res, nb)) > 0).astype(int)
    assert np.array_equal(dec, bits[0] & bits[1])
    L.bootsXOR_fullGPU_n_Bit(res, sa, sb, nb, handle, None, None)
    assert np.array_equal((o.phases(keys.lwe_key, read(res, nb)) > 0).astype(int), bits[0] ^ bits[1])
    L.bootsMUX_fullGPU_n_Bit(res, sa, sb, sc, nb, handle, None, None)
    assert np.array_equal((o.phases(keys.lwe_key, read(res, nb)) > 0).astype(int),
                          np.where(bits[0] == 1, bits[1], bits[2]))
    L.bootsANDXOR_fullGPU_n_Bit_vector(res, sa, sb, 1, nb, handle, None, None)
    dec = (o.phases(keys.lwe_key, read(res, 2 * nb)) > 0).astype(int)
    assert np.array_equal(dec[:nb], bits[0] & bits[1]) and np.array_equal(dec[nb:], bits[0] ^ bits[1])
    L.bootsXORXOR_fullGPU_n_Bit_vector(res, sa, sb, sc, sb, 1, nb, handle, None, None)
    dec = (o.phases(keys.lwe_key, read(res, 2 * nb)) > 0).astype(int)
    assert np.array_equal(dec[:nb], bits[0] ^ bits[1]) and np.array_equal(dec[nb:], bits[2] ^ bits[1])
    for s in (sa, sb, sc, res):
        L.freeLweSample_16_gpu(s)


def test_cloud_program_flow_through_files(pkg, tmp_path):
    """cpu/cloud.cpp:138-161 + its gate loop, with THIS library's reference-named entry points only:
    read cloud.key, read cloud.data, evaluate gates, write answer.data; the client (flat API)
    wrote the inputs and decrypts the answer."""
    sk = pkg.keygen(21)
    bits_a = np.array([0, 0, 1, 1, 1, 0], np.int32)
    bits_b = np.array([0, 1, 0, 1, 1, 1], np.int32)
    pkg.write_cloud_key(tmp_path / "cloud.key", sk)
    pkg.write_ciphertexts(tmp_path / "cloud.data", pkg.encrypt_bits(sk, bits_a, 1))
    pkg.write_ciphertexts(tmp_path / "cloud.data", pkg.encrypt_bits(sk, bits_b, 2), append=True)

    L = ctypes.CDLL(pkg.lib_path())
    libc = ctypes.CDLL(None)
    libc.fopen.restype = vp
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [vp]
    L.new_tfheGateBootstrappingCloudKeySet_fromFile.restype = vp
    L.new_tfheGateBootstrappingCloudKeySet_fromFile.argtypes = [vp]
    L.new_gate_bootstrapping_ciphertext_array.restype = vp
    L.new_gate_bootstrapping_ciphertext_array.argtypes = [ctypes.c_int, vp]
    for f in ("import_gate_bootstrapping_ciphertext_fromFile", "export_gate_bootstrapping_ciphertext_toFile"):
        getattr(L, f).argtypes = [vp, vp, vp]
    for f in ("bootsAND", "bootsXOR", "bootsNAND"):
        getattr(L, f).argtypes = [vp, vp, vp, vp]
    L.bootsMUX.argtypes = [vp, vp, vp, vp, vp]

    F = libc.fopen(str(tmp_path / "cloud.key").encode(), b"rb")
    bk = vp(L.new_tfheGateBootstrappingCloudKeySet_fromFile(F))
    libc.fclose(F)
    params = ctypes.cast(bk, ctypes.POINTER(vp))[0]  # TFheGateBootstrappingCloudKeySet::params (first member)
    n = len(bits_a)
    SZ = 24  # sizeof(LweSample): pointer, int32 (+pad), double
    a = L.new_gate_bootstrapping_ciphertext_array(n, params)
    b = L.new_gate_bootstrapping_ciphertext_array(n, params)
    out = L.new_gate_bootstrapping_ciphertext_array(4 * n, params)
    F = libc.fopen(str(tmp_path / "cloud.data").encode(), b"rb")
    for arr in (a, b):
        for i in range(n):
            L.import_gate_bootstrapping_ciphertext_fromFile(F, arr + i * SZ, params)
    libc.fclose(F)
    for i in range(n):
        L.bootsAND(out + i * SZ, a + i * SZ, b + i * SZ, bk)
        L.bootsXOR(out + (n + i) * SZ, a + i * SZ, b + i * SZ, bk)
        L.bootsNAND(out + (2 * n + i) * SZ, a + i * SZ, b + i * SZ, bk)
        L.bootsMUX(out + (3 * n + i) * SZ, a + i * SZ, b + i * SZ, a + ((i + 1) % n) * SZ, bk)
    F = libc.fopen(str(tmp_path / "answer.data").encode(), b"wb")
    for i in range(4 * n):
        L.export_gate_bootstrapping_ciphertext_toFile(F, out + i * SZ, params)
    libc.fclose(F)
    # re-export of the key set reproduces the file
    F = libc.fopen(str(tmp_path / "cloud2.key").encode(), b"wb")
    L.export_tfheGateBootstrappingCloudKeySet_toFile.argtypes = [vp, vp]
    L.export_tfheGateBootstrappingCloudKeySet_toFile(F, bk)
    libc.fclose(F)
    import filecmp
    assert filecmp.cmp(tmp_path / "cloud.key", tmp_path / "cloud2.key", shallow=False)
    for arr, cnt in ((a, n), (b, n), (out, 4 * n)):
        L.delete_gate_bootstrapping_ciphertext_array.argtypes = [ctypes.c_int, vp]
        L.delete_gate_bootstrapping_ciphertext_array(cnt, arr)
    L.tfhe_b200_delete_cloud_keyset_fromFile.argtypes = [vp]
    L.tfhe_b200_delete_cloud_keyset_fromFile(bk)

    ans, _ = pkg.read_ciphertexts(tmp_path / "answer.data", 500)
    got = pkg.decrypt_bits(sk, ans)
    mux = np.where(bits_a == 1, bits_b, np.roll(bits_a, -1))
    assert np.array_equal(got, np.concatenate([bits_a & bits_b, bits_a ^ bits_b, 1 - (bits_a & bits_b), mux]))


def test_classic_gates_coalesce_across_threads(world):
    """SURVEY 8b "Threading": the classic entry points are called from OpenMP workers
    (cpuParallel/Cipher.cpp:75-76, 94).  16 host threads x bootsAND / bootsXOR / bootsMUX on reference
    objects: every result decrypts correctly, the calls shared launches (gate coalescer), and the
    16 calls together take less than twice one lone call."""
    import threading
    import time

    o, r, R, L, keys = world
    cloud = vp(R.ref_cloud_keyset(r.h))
    nthreads = 16
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, (nthreads, 3))
    samples = [[Sample(R, r.h, 500, r.encrypt(int(b))) for b in row] for row in bits]
    results = [Sample(R, r.h, 500) for _ in range(nthreads)]
    kinds = ["AND", "XOR", "MUX", "NAND"]

    def call(i):
        k = kinds[i % len(kinds)]
        a, b, c = samples[i]
        if k == "MUX":
            L.bootsMUX(results[i].p, a.p, b.p, c.p, cloud)
        else:
            getattr(L, "boots" + k)(results[i].p, a.p, b.p, cloud)

    def expected(i):
        a, b, c = (int(x) for x in bits[i])
        return {"AND": a & b, "XOR": a ^ b, "MUX": b if a else c, "NAND": 1 - (a & b)}[kinds[i % len(kinds)]]

    call(0)  # context creation + key upload happen here
    lone = []
    for _ in range(3):
        t0 = time.perf_counter()
        call(0)
        lone.append(time.perf_counter() - t0)
    b0, g0 = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    L.tfhe_b200_compat_coalescer_stats(cloud, ctypes.byref(b0), ctypes.byref(g0))
    best = None
    for _ in range(3):
        threads = [threading.Thread(target=call, args=(i,)) for i in range(nthreads)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    b1, g1 = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    L.tfhe_b200_compat_coalescer_stats(cloud, ctypes.byref(b1), ctypes.byref(g1))
    for i in range(nthreads):
        assert int(r.phase(results[i].get()) > 0) == expected(i), i
    assert g1.value - g0.value == 3 * nthreads
    assert b1.value - b0.value < 3 * nthreads / 2, "calls from different threads did not share launches"
    assert best < 3.0 * min(lone) + 2e-3, (best, lone)   # Python threads start ~0.1 ms apart: two rounds

    # the same from real OpenMP workers (cpuParallel/Cipher.cpp:114-121 cipherAND): 16 threads x bootsAND
    # arrive together, share ONE launch and finish in less than twice a lone call
    R.ref_omp_gate_calls.restype = ctypes.c_double
    R.ref_omp_gate_calls.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]
    fn = ctypes.cast(L.bootsAND, vp)
    arr = lambda objs: (vp * nthreads)(*[x.p for x in objs])
    res_p, a_p, b_p = arr(results), arr([s[0] for s in samples]), arr([s[1] for s in samples])
    R.ref_omp_gate_calls(fn, res_p, a_p, b_p, cloud, nthreads, nthreads)  # thread pool start-up
    L.tfhe_b200_compat_coalescer_stats(cloud, ctypes.byref(b0), ctypes.byref(g0))
    omp = min(R.ref_omp_gate_calls(fn, res_p, a_p, b_p, cloud, nthreads, nthreads) for _ in range(5))
    L.tfhe_b200_compat_coalescer_stats(cloud, ctypes.byref(b1), ctypes.byref(g1))
    for i in range(nthreads):
        assert int(r.phase(results[i].get()) > 0) == int(bits[i][0] & bits[i][1]), i
    assert g1.value - g0.value == 5 * nthreads and b1.value - b0.value <= 10
    assert omp < 2.0 * min(lone), (omp, lone)


def test_xnor_not16_and_keys_free(world):
    """The remaining exports of the batched family: bootsXNOR_fullGPU_n_Bit, bootsNOT_16
    (boot-gates.cu:2953, :1274) and tfhe_b200_keys_free (contexts are rebuilt on the next use)."""
    import torch

    o, r, R, L, keys = world

    class S16(ctypes.Structure):
        _fields_ = [("a", vp), ("b", ctypes.POINTER(ctypes.c_int)), ("cv", ctypes.POINTER(ctypes.c_double))]

    cloud = vp(R.ref_cloud_keyset(r.h))
    handle = vp(L.tfhe_b200_keys_to_gpu(cloud))
    nb = 5
    rng = np.random.default_rng(6)
    bits = [rng.integers(0, 2, nb) for _ in range(2)]
    torch.zeros(1, device="cuda")
    cudart = ctypes.CDLL("libcudart.so.12")
    cudart.cudaMemcpy.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_int]

    def make(bitvec):
        s = ctypes.cast(L.convertBitToNumberZero_GPU(nb, cloud), ctypes.POINTER(S16))
        flat = np.stack([r.encrypt(int(b)) for b in bitvec])
        a = np.ascontiguousarray(flat[:, :-1])
        assert cudart.cudaMemcpy(s.contents.a, a.ctypes.data_as(vp), a.nbytes, 1) == 0
        for i in range(nb):
            s.contents.b[i] = int(flat[i, -1])
        return s, flat

    def read(s, count):
        a = np.zeros((count, 500), np.int32)
        assert cudart.cudaMemcpy(a.ctypes.data_as(vp), s.contents.a, a.nbytes, 2) == 0
        b = np.array([s.contents.b[i] for i in range(count)], np.int32)
        return np.concatenate([a, b[:, None]], 1)

    sa, fa = make(bits[0])
    sb, fb = make(bits[1])
    res = ctypes.cast(L.convertBitToNumberZero_GPU(nb, cloud), ctypes.POINTER(S16))
    L.bootsXNOR_fullGPU_n_Bit(res, sa, sb, nb, handle, None, None)
    assert np.array_equal((o.phases(keys.lwe_key, read(res, nb)) > 0).astype(int), 1 - (bits[0] ^ bits[1]))
    L.bootsNOT_16(res, sa, nb, 500)
    assert np.array_equal(read(res, nb), (-fa.astype(np.int64)).astype(np.int32))  # bit-exact negation
    # keys_free drops the context; the next use rebuilds it and still computes correctly
    L.tfhe_b200_compat_cached_contexts.restype = ctypes.c_int
    before = L.tfhe_b200_compat_cached_contexts()
    L.tfhe_b200_keys_free(cloud)
    assert L.tfhe_b200_compat_cached_contexts() == before - 1
    handle = vp(L.tfhe_b200_keys_to_gpu(cloud))
    L.bootsAND_fullGPU_n_Bit(res, sa, sb, nb, handle, None, None)
    assert np.array_equal((o.phases(keys.lwe_key, read(res, nb)) > 0).astype(int), bits[0] & bits[1])
    for s in (sa, sb, res):
        L.freeLweSample_16_gpu(s)


def test_context_cache_follows_key_content(world):
    """A GPU context is cached per host key object; the cache is validated by a fingerprint of the
    key material, so a TGSW sample OVERWRITTEN IN PLACE (same address, new content) is re-uploaded,
    and the number of bare-TGSW contexts stays bounded."""
    o, r, R, L, keys = world
    rng = np.random.default_rng(18)
    acc = rng.integers(-2 ** 31, 2 ** 31, (2, 1024), dtype=np.int64).astype(np.int32)
    params = vp(R.ref_tgsw_params(r.h))
    base = R.ref_tgsw_fft_array(r.h)
    t = vp(R.ref_tlwe_new(r.h))
    got = np.zeros_like(acc)
    L.tfhe_b200_compat_cached_contexts.restype = ctypes.c_int

    def extern_mul(ptr):
        R.ref_tlwe_set(t, acc.ctypes.data_as(vp), 1024, 1)
        L.tGswFFTExternMulToTLwe(t, vp(ptr), params)
        R.ref_tlwe_get(t, got.ctypes.data_as(vp), 1024, 1)
        return got.copy()

    # copy TGSW sample 3 into a scratch object at a fixed address, use it, then overwrite it with sample 9
    # (struct TGswSampleFFT {all_samples, sample, k, l}: 24 bytes; data = 4 rows x 2 polys x 512 complex)
    class LagrangeHalfC(ctypes.Structure):
        _fields_ = [("data", vp), ("precomp", vp)]

    class TLweSampleFFT(ctypes.Structure):
        _fields_ = [("a", ctypes.POINTER(LagrangeHalfC)), ("b", vp), ("cv", ctypes.c_double), ("k", ctypes.c_int)]

    class TGswSampleFFT(ctypes.Structure):
        _fields_ = [("all_samples", ctypes.POINTER(TLweSampleFFT)), ("sample", vp), ("k", ctypes.c_int), ("l", ctypes.c_int)]

    def poly_ptrs(idx):
        s = ctypes.cast(vp(base + 24 * idx), ctypes.POINTER(TGswSampleFFT)).contents
        return [s.all_samples[row].a[j].data for row in range(4) for j in range(2)]

    scratch_idx, src_a, src_b = 20, 3, 9
    saved = [ctypes.string_at(p, 8192) for p in poly_ptrs(scratch_idx)]
    try:
        for dst, src in zip(poly_ptrs(scratch_idx), poly_ptrs(src_a)):
            ctypes.memmove(dst, src, 8192)
        first = extern_mul(base + 24 * scratch_idx)
        assert np.abs(wrap32(first.astype(np.int64) - o.extern_mul_exact(keys.bk[src_a], acc).astype(np.int64))).max() <= 1
        for dst, src in zip(poly_ptrs(scratch_idx), poly_ptrs(src_b)):
            ctypes.memmove(dst, src, 8192)   # same object, same address, new key material
        second = extern_mul(base + 24 * scratch_idx)
        assert np.abs(wrap32(second.astype(np.int64) - o.extern_mul_exact(keys.bk[src_b], acc).astype(np.int64))).max() <= 1
    finally:
        for dst, data in zip(poly_ptrs(scratch_idx), saved):
            ctypes.memmove(dst, data, 8192)
    # many distinct TGSW objects: the cache keeps at most 8 bare-TGSW contexts
    L.tfhe_b200_compat_release_all()
    for idx in range(40, 52):
        extern_mul(base + 24 * idx)
    assert L.tfhe_b200_compat_cached_contexts() <= 8
    L.tfhe_b200_compat_invalidate(vp(base + 24 * 51))
    assert L.tfhe_b200_compat_cached_contexts() <= 7
    R.ref_tlwe_free(t)
