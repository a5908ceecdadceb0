"""ctypes binding of include/tfhe_b200.h and a thin tensor-level wrapper.

Function names and argument meaning mirror the reference's operator interface
for the path (gpuParallel/tfhe_gate_bootstrapping_functions.h:14-198):
``Engine.gate("NAND", ca, cb)`` is the batched ``bootsNAND`` and so on; errors
raise ``EngineError`` (the reference aborts the process instead,
tfhe_gate_bootstrapping.cu:11-15).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GATES = ["NAND", "OR", "AND", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]
GATE_ID = {g: i for i, g in enumerate(GATES)}
MU = 0x20000000  # modSwitchToTorus32(1, 8)

_vp = ctypes.c_void_p
_i = ctypes.c_int


CMP = {"GT": 0, "LE": 1, "LT": 2, "GE": 3, "EQ": 4, "NE": 5}   # TFHE_B200_CMP_*
SHIFT = {"LEFT": 0, "RIGHT_LOGICAL": 1, "RIGHT_ARITH": 2}         # TFHE_B200_SHIFT_*
GPC = 10      # TFHE_B200_GPC: carry operator g | (p & c) for mutually exclusive g, p (extension)
XOR3, MAJ = 11, 12   # full-adder outputs a ^ b ^ c and majority(a, b, c), one bootstrap each (extensions)
SUMC = 13            # a ^ (b | (c & d)), b and c exclusive: fused sum bit of the prefix adder (extension)
ADDER = {"RIPPLE": 0, "PREFIX": 1, "CARRY_SAVE": 2}   # TFHE_B200_ADDER_*


class EngineError(RuntimeError):
    pass


class GateOp(ctypes.Structure):
    """tfhe_b200_gate_op (include/tfhe_b200.h)."""
    _fields_ = [("gate", ctypes.c_int32), ("count", ctypes.c_int32), ("a", ctypes.c_void_p), ("b", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("stride_a", ctypes.c_int64), ("stride_b", ctypes.c_int64),
                ("stride_out", ctypes.c_int64), ("idx_a", ctypes.c_void_p), ("idx_b", ctypes.c_void_p),
                ("idx_out", ctypes.c_void_p), ("c", ctypes.c_void_p), ("stride_c", ctypes.c_int64),
                ("idx_c", ctypes.c_void_p), ("d", ctypes.c_void_p), ("stride_d", ctypes.c_int64),
                ("idx_d", ctypes.c_void_p)]


class Params(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n", "N", "k", "l", "Bgbit", "ks_t", "ks_basebit")]


def lib_path():
    # TFHE_B200_LIB: development override used to A/B differently compiled builds of the SAME
    # library on the GPU box (tools/build_variant.py); there is no other implementation to fall back to
    return os.environ.get("TFHE_B200_LIB") or os.path.join(_HERE, "libtfhe_b200.so")


_lib = None


def lib():
    """Loads libtfhe_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise EngineError("%s not built: run `python cpu-gpu-tfhe_b200/build.py`" % p)
        L = ctypes.CDLL(p)
        L.tfhe_b200_last_error.restype = ctypes.c_char_p
        L.tfhe_b200_key_bytes.restype = ctypes.c_size_t
        L.tfhe_b200_launch_count.restype = ctypes.c_ulonglong
        L.tfhe_b200_key_bytes.argtypes = [_vp]
        L.tfhe_b200_launch_count.argtypes = [_vp]
        L.tfhe_b200_sm_count.argtypes = [_vp]
        L.tfhe_b200_ctx_create.argtypes = [ctypes.POINTER(_vp), ctypes.POINTER(Params), _i]
        L.tfhe_b200_ctx_destroy.argtypes = [_vp]
        L.tfhe_b200_load_keys.argtypes = [_vp, _vp, _vp]
        L.tfhe_b200_load_keys_device.argtypes = [_vp, _vp, _vp, _vp]
        L.tfhe_b200_load_bk_fourier.argtypes = [_vp, _vp]
        L.tfhe_b200_load_ks.argtypes = [_vp, _vp]
        L.tfhe_b200_gate.argtypes = [_vp, _i, _vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_gate2.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_gate_pair.argtypes = [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_mux.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_gate_multi.argtypes = [_vp, _vp, _i, _vp]
        L.tfhe_b200_not.argtypes = [_vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_copy.argtypes = [_vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_constant.argtypes = [_vp, _vp, _i, _i, _vp]
        L.tfhe_b200_bootstrap_woks.argtypes = [_vp, _vp, _vp, ctypes.c_int32, _i, _vp]
        L.tfhe_b200_bootstrap.argtypes = [_vp, _vp, _vp, ctypes.c_int32, _i, _vp]
        L.tfhe_b200_keyswitch.argtypes = [_vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_blind_rotate.argtypes = [_vp, _vp, _vp, _i, _i, _vp]
        L.tfhe_b200_blind_rotate_and_extract.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]
        L.tfhe_b200_extern_mul.argtypes = [_vp, _vp, _i, _i, _vp]
        L.tfhe_b200_gate_host.argtypes = [_vp, _i, _vp, _vp, _vp, _i]
        L.tfhe_b200_mux_host.argtypes = [_vp, _vp, _vp, _vp, _vp, _i]
        L.tfhe_b200_bk_words.restype = ctypes.c_size_t
        L.tfhe_b200_ks_words.restype = ctypes.c_size_t
        L.tfhe_b200_bk_words.argtypes = [ctypes.POINTER(Params)]
        L.tfhe_b200_ks_words.argtypes = [ctypes.POINTER(Params)]
        L.tfhe_b200_keygen.argtypes = [ctypes.POINTER(Params), ctypes.c_uint64, ctypes.c_double, ctypes.c_double,
                                       _vp, _vp, _vp, _vp]
        L.tfhe_b200_encrypt_bits.argtypes = [ctypes.POINTER(Params), _vp, ctypes.c_uint64, ctypes.c_double, _vp, _i,
                                             _vp]
        L.tfhe_b200_decrypt_bits.argtypes = [ctypes.POINTER(Params), _vp, _vp, _i, _vp]
        L.tfhe_b200_phases.argtypes = [_vp, _i, _vp, _i, _vp]
        L.tfhe_b200_set_timing.argtypes = [_vp, _i]
        for f in ("add", "mul", "matmul", "mul_ex", "matmul_ex", "sub", "neg", "compare", "minmax", "select", "abs",
                  "shift", "div", "mul_full", "mul_karatsuba", "matmul_cannon"):
            getattr(L, "tfhe_b200_circuit_" + f).restype = _vp
        L.tfhe_b200_circuit_mul_full.argtypes = [_vp, _i, _i, _i]
        L.tfhe_b200_circuit_mul_karatsuba.argtypes = [_vp, _i, _i, _i]
        L.tfhe_b200_circuit_matmul_cannon.argtypes = [_vp, _i, _i, _i]
        L.tfhe_b200_circuit_sub.argtypes = [_vp, _i, _i, _i]
        L.tfhe_b200_circuit_neg.argtypes = [_vp, _i, _i]
        L.tfhe_b200_circuit_compare.argtypes = [_vp, _i, _i, _i, _i]
        L.tfhe_b200_circuit_minmax.argtypes = [_vp, _i, _i, _i, _i]
        L.tfhe_b200_circuit_select.argtypes = [_vp, _i, _i]
        L.tfhe_b200_circuit_abs.argtypes = [_vp, _i, _i, _i]
        L.tfhe_b200_circuit_shift.argtypes = [_vp, _i, _i, _i, _i]
        L.tfhe_b200_circuit_div.argtypes = [_vp, _i, _i, _i, _i]
        L.tfhe_b200_mux_gather.argtypes = [_vp, _vp, ctypes.c_int64, _vp, _vp, _vp, _vp, _i, _vp]
        L.tfhe_b200_linear_gather.argtypes = [_vp, _vp, ctypes.c_int64, _vp, _vp, _i, ctypes.c_int32, _i, _vp]
        L.tfhe_b200_circuit_add.argtypes = [_vp, _i, _i, _i]
        L.tfhe_b200_circuit_mul.argtypes = [_vp, _i, _i]
        L.tfhe_b200_circuit_mul_ex.argtypes = [_vp, _i, _i, _i]
        L.tfhe_b200_circuit_matmul.argtypes = [_vp, _i, _i, _i, _i]
        L.tfhe_b200_circuit_matmul_ex.argtypes = [_vp, _i, _i, _i, _i, _i]
        L.tfhe_b200_circuit_simulate.argtypes = [_vp, _vp, _vp]
        L.tfhe_b200_circuit_destroy.argtypes = [_vp]
        L.tfhe_b200_circuit_levels.argtypes = [_vp]
        L.tfhe_b200_circuit_gates.argtypes = [_vp]
        L.tfhe_b200_circuit_gates.restype = ctypes.c_longlong
        L.tfhe_b200_circuit_level_gates.argtypes = [_vp, ctypes.c_int]
        L.tfhe_b200_circuit_level_gates.restype = ctypes.c_longlong
        L.tfhe_b200_circuit_operands.argtypes = [_vp]
        L.tfhe_b200_circuit_operand_rows.argtypes = [_vp, _i]
        L.tfhe_b200_circuit_output_rows.argtypes = [_vp]
        L.tfhe_b200_circuit_run.argtypes = [_vp, _vp, _vp, _vp]
        L.tfhe_b200_circuit_run_many.argtypes = [_vp, _i, _vp, _vp, _vp]
        L.tfhe_b200_circuit_set_graph.argtypes = [_vp, _i]
        L.tfhe_b200_circuit_used_graph.argtypes = [_vp]
        L.tfhe_b200_count_launches.argtypes = [_vp, ctypes.c_ulonglong]
        L.tfhe_b200_file_read_cloud_key.argtypes = [ctypes.c_char_p, _vp, _vp, _vp, _vp, _vp]
        L.tfhe_b200_file_write_cloud_key.argtypes = [ctypes.c_char_p, _vp, _vp, _vp, _vp, _vp]
        L.tfhe_b200_file_read_secret_key.argtypes = [ctypes.c_char_p, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.tfhe_b200_file_write_secret_key.argtypes = [ctypes.c_char_p, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.tfhe_b200_file_count_ciphertexts.argtypes = [ctypes.c_char_p, _i]
        L.tfhe_b200_file_count_ciphertexts.restype = ctypes.c_long
        L.tfhe_b200_file_read_ciphertexts.argtypes = [ctypes.c_char_p, _i, _vp, _vp, _i]
        L.tfhe_b200_file_write_ciphertexts.argtypes = [ctypes.c_char_p, _i, _vp, _vp, _i, _i]
        L.tfhe_b200_file_last_error.restype = ctypes.c_char_p
        L.tfhe_b200_keygen_device.argtypes = [_vp, _vp, ctypes.c_uint64, ctypes.c_double, ctypes.c_double, _vp, _vp,
                                              _vp, _vp]
        L.tfhe_b200_get_timing.argtypes = [_vp, _vp, _vp, _vp]
        L.tfhe_b200_measure_fp64_peak.argtypes = [_i, _vp, _vp]
        L.tfhe_b200_multi_create.argtypes = [ctypes.POINTER(_vp), ctypes.POINTER(Params), _vp, _i]
        L.tfhe_b200_multi_destroy.argtypes = [_vp]
        L.tfhe_b200_multi_devices.argtypes = [_vp]
        L.tfhe_b200_multi_ctx.argtypes = [_vp, _i]
        L.tfhe_b200_multi_ctx.restype = _vp
        L.tfhe_b200_multi_load_keys.argtypes = [_vp, _vp, _vp]
        L.tfhe_b200_multi_gate_host.argtypes = [_vp, _i, _vp, _vp, _vp, ctypes.c_longlong]
        L.tfhe_b200_multi_mux_host.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.c_longlong]
        L.tfhe_b200_multi_launch_count.argtypes = [_vp]
        L.tfhe_b200_multi_launch_count.restype = ctypes.c_ulonglong
        L.tfhe_b200_multi_last_error.restype = ctypes.c_char_p
        L.tfhe_b200_conversion_mode.restype = _i
        _lib = L
    return _lib


class SecretKeys:
    """Client-side key material in the flat formats of include/tfhe_b200.h."""

    def __init__(self, params, lwe_key, tlwe_key, bk, ks, alpha_lwe, alpha_bk):
        self.params, self.lwe_key, self.tlwe_key, self.bk, self.ks = params, lwe_key, tlwe_key, bk, ks
        self.alpha_lwe, self.alpha_bk = alpha_lwe, alpha_bk


def keygen(seed, params=None):
    """new_random_gate_bootstrapping_secret_keyset (tfhe_gate_bootstrapping.cu:57-68), host side."""
    L = lib()
    p = params or default_params()
    a, b = ctypes.c_double(), ctypes.c_double()
    L.tfhe_b200_default_noise(ctypes.byref(a), ctypes.byref(b))
    kpl = (p.k + 1) * p.l
    lwe = np.zeros(p.n, np.int32)
    tlwe = np.zeros(p.k * p.N, np.int32)
    bk = np.zeros((p.n, kpl, p.k + 1, p.N), np.int32)
    ks = np.zeros((p.N * p.k, p.ks_t, 1 << p.ks_basebit, p.n + 1), np.int32)
    if L.tfhe_b200_keygen(ctypes.byref(p), seed, a.value, b.value, lwe.ctypes.data, tlwe.ctypes.data,
                          bk.ctypes.data, ks.ctypes.data):
        raise EngineError("keygen failed")
    return SecretKeys(p, lwe, tlwe, bk, ks, a.value, b.value)


# ---- key / ciphertext files of stock TFHE clients (tfhe_io.cu formats) ------------------------

def _file_ck(rc):
    if rc:
        raise EngineError(lib().tfhe_b200_file_last_error().decode())


def _alphas(sk):
    # new_default_gate_bootstrapping_parameters (tfhe_gate_bootstrapping.cu:36-38):
    # max_stdev = sqrt(2/pi) * 2^-4 / 4 for both parameter sets
    amax = (2.0 / np.pi) ** 0.5 * 2.0 ** -6
    return (ctypes.c_double * 4)(sk.alpha_lwe, amax, sk.alpha_bk, amax)


def write_cloud_key(path, sk, variances=None):
    """cloud.key (export_tfheGateBootstrappingCloudKeySet_toFile, tfhe_io.cu:1109)."""
    v = (ctypes.c_double * 2)(*(variances if variances is not None else (sk.alpha_bk ** 2, sk.alpha_lwe ** 2)))
    _file_ck(lib().tfhe_b200_file_write_cloud_key(str(path).encode(), ctypes.byref(sk.params), _alphas(sk), v,
                                                  sk.bk.ctypes.data, sk.ks.ctypes.data))


def write_secret_key(path, sk, variances=None):
    """secret.key (export_tfheGateBootstrappingSecretKeySet_toFile, tfhe_io.cu:1173)."""
    v = (ctypes.c_double * 2)(*(variances if variances is not None else (sk.alpha_bk ** 2, sk.alpha_lwe ** 2)))
    _file_ck(lib().tfhe_b200_file_write_secret_key(str(path).encode(), ctypes.byref(sk.params), _alphas(sk), v,
                                                   sk.bk.ctypes.data, sk.ks.ctypes.data, sk.lwe_key.ctypes.data,
                                                   sk.tlwe_key.ctypes.data))


def read_key_file(path, secret=False):
    """Reads cloud.key / secret.key; returns (SecretKeys [lwe_key / tlwe_key None for a cloud key], variances)."""
    L = lib()
    p = Params()
    al, var = (ctypes.c_double * 4)(), (ctypes.c_double * 2)()
    pb = str(path).encode()
    _file_ck(L.tfhe_b200_file_read_cloud_key(pb, ctypes.byref(p), al, None, None, None))  # header
    kpl = (p.k + 1) * p.l
    bk = np.zeros((p.n, kpl, p.k + 1, p.N), np.int32)
    ks = np.zeros((p.N * p.k, p.ks_t, 1 << p.ks_basebit, p.n + 1), np.int32)
    lwe = tlwe = None
    if secret:
        lwe, tlwe = np.zeros(p.n, np.int32), np.zeros(p.k * p.N, np.int32)
        _file_ck(L.tfhe_b200_file_read_secret_key(pb, ctypes.byref(p), al, var, bk.ctypes.data, ks.ctypes.data,
                                                  lwe.ctypes.data, tlwe.ctypes.data))
    else:
        _file_ck(L.tfhe_b200_file_read_cloud_key(pb, ctypes.byref(p), al, var, bk.ctypes.data, ks.ctypes.data))
    return SecretKeys(p, lwe, tlwe, bk, ks, al[0], al[2]), list(var)


def write_ciphertexts(path, samples, variances=None, append=False):
    """cloud.data / answer.data: one record per sample (export_gate_bootstrapping_ciphertext_toFile)."""
    s = _np_i32(samples)
    n = s.shape[-1] - 1
    flat = s.reshape(-1, n + 1)
    v = None if variances is None else np.ascontiguousarray(variances, np.float64)
    _file_ck(lib().tfhe_b200_file_write_ciphertexts(str(path).encode(), n, flat.ctypes.data,
                                                    None if v is None else v.ctypes.data, flat.shape[0],
                                                    1 if append else 0))


def read_ciphertexts(path, n, count=None):
    L = lib()
    if count is None:
        count = int(L.tfhe_b200_file_count_ciphertexts(str(path).encode(), n))
        if count < 0:
            raise EngineError("%s is not a whole number of ciphertext records" % path)
    out = np.zeros((count, n + 1), np.int32)
    var = np.zeros(count, np.float64)
    _file_ck(L.tfhe_b200_file_read_ciphertexts(str(path).encode(), n, out.ctypes.data, var.ctypes.data, count))
    return out, var


def encrypt_bits(sk, bits, seed):
    """bootsSymEncrypt (tfhe_gate_bootstrapping.cu:114) for an array of bits."""
    bits = _np_i32(bits).reshape(-1)
    out = np.empty((bits.size, sk.params.n + 1), np.int32)
    if lib().tfhe_b200_encrypt_bits(ctypes.byref(sk.params), sk.lwe_key.ctypes.data, seed, sk.alpha_lwe,
                                    bits.ctypes.data, bits.size, out.ctypes.data):
        raise EngineError("encrypt failed")
    return out


def decrypt_bits(sk, samples):
    """bootsSymDecrypt (tfhe_gate_bootstrapping.cu:122)."""
    samples = _np_i32(samples)
    flat = samples.reshape(-1, sk.params.n + 1)
    out = np.empty(flat.shape[0], np.int32)
    if lib().tfhe_b200_decrypt_bits(ctypes.byref(sk.params), sk.lwe_key.ctypes.data, flat.ctypes.data,
                                    flat.shape[0], out.ctypes.data):
        raise EngineError("decrypt failed")
    return out.reshape(samples.shape[:-1])


def phases(key, samples):
    samples = _np_i32(samples)
    key = _np_i32(key)
    flat = samples.reshape(-1, key.size + 1)
    out = np.empty(flat.shape[0], np.int32)
    lib().tfhe_b200_phases(key.ctypes.data, key.size, flat.ctypes.data, flat.shape[0], out.ctypes.data)
    return out.reshape(samples.shape[:-1])


def measure_fp64_peak(device=0):
    """(burst, sustained) fp64 FMA TFLOP/s measured live on `device`."""
    a, b = ctypes.c_double(), ctypes.c_double()
    if lib().tfhe_b200_measure_fp64_peak(device, ctypes.byref(a), ctypes.byref(b)):
        raise EngineError("fp64 peak measurement failed")
    return a.value, b.value


def default_params():
    p = Params()
    lib().tfhe_b200_default_params(ctypes.byref(p))
    return p


def device_count():
    return int(lib().tfhe_b200_device_count())


def _np_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    """One engine context = one GPU + one key set."""

    def __init__(self, params=None, device=0):
        self.L = lib()
        self.p = params or default_params()
        self.device = device
        h = _vp()
        if self.L.tfhe_b200_ctx_create(ctypes.byref(h), ctypes.byref(self.p), device):
            raise EngineError(self.L.tfhe_b200_last_error().decode())
        self.h = h
        self.words = self.p.n + 1

    def close(self):
        if getattr(self, "h", None):
            self.L.tfhe_b200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise EngineError(self.L.tfhe_b200_last_error().decode())

    # -- keys ----------------------------------------------------------------
    def load_keys(self, bk_coef=None, ks=None):
        bk = _np_i32(bk_coef) if bk_coef is not None else None
        k = _np_i32(ks) if ks is not None else None
        self._ck(self.L.tfhe_b200_load_keys(self.h, bk.ctypes.data if bk is not None else None,
                                            k.ctypes.data if k is not None else None))

    def load_keys_device(self, d_bk_coef=None, d_ks=None, stream=None):
        self._ck(self.L.tfhe_b200_load_keys_device(
            self.h, d_bk_coef.data_ptr() if d_bk_coef is not None else None,
            d_ks.data_ptr() if d_ks is not None else None, self._stream(stream)))

    def load_bk_fourier(self, bkfft_ref):
        a = np.ascontiguousarray(bkfft_ref, dtype=np.complex128)
        self._ck(self.L.tfhe_b200_load_bk_fourier(self.h, a.ctypes.data))

    def keygen(self, seed, want_flat=True):
        """Fresh keys generated on this engine's GPU and loaded into it (tfhe_b200_keygen_device).
        Returns SecretKeys (bk / ks are None unless want_flat)."""
        p = self.p
        a, b = ctypes.c_double(), ctypes.c_double()
        self.L.tfhe_b200_default_noise(ctypes.byref(a), ctypes.byref(b))
        kpl = (p.k + 1) * p.l
        lwe, tlwe = np.zeros(p.n, np.int32), np.zeros(p.k * p.N, np.int32)
        bk = np.zeros((p.n, kpl, p.k + 1, p.N), np.int32) if want_flat else None
        ks = np.zeros((p.N * p.k, p.ks_t, 1 << p.ks_basebit, p.n + 1), np.int32) if want_flat else None
        self._ck(self.L.tfhe_b200_keygen_device(self.h, ctypes.byref(p), seed, a.value, b.value, lwe.ctypes.data,
                                                tlwe.ctypes.data, bk.ctypes.data if want_flat else None,
                                                ks.ctypes.data if want_flat else None))
        return SecretKeys(p, lwe, tlwe, bk, ks, a.value, b.value)

    @property
    def key_bytes(self):
        return int(self.L.tfhe_b200_key_bytes(self.h))

    @property
    def launch_count(self):
        return int(self.L.tfhe_b200_launch_count(self.h))

    @property
    def sm_count(self):
        return int(self.L.tfhe_b200_sm_count(self.h))

    def set_timing(self, enable):
        self._ck(self.L.tfhe_b200_set_timing(self.h, int(bool(enable))))

    def get_timing(self):
        """(blind_rotate_ms, keyswitch_ms, calls) accumulated since set_timing(True)."""
        a, b, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        self._ck(self.L.tfhe_b200_get_timing(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
        return a.value, b.value, n.value

    # -- helpers ---------------------------------------------------------------
    def _stream(self, stream):
        import torch

        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return _vp(s.cuda_stream)

    def _dev(self):
        import torch

        return torch.device("cuda", self.device)

    def _chk(self, t, words=None):
        import torch

        assert t.is_cuda and t.dtype == torch.int32 and t.is_contiguous(), "need contiguous int32 CUDA tensor"
        if words is not None:
            assert t.shape[-1] == words, "last dimension must be %d" % words
        return t

    def empty(self, count, words=None):
        import torch

        return torch.empty((count, words or self.words), dtype=torch.int32, device=self._dev())

    def to_device(self, a):
        import torch

        return torch.from_numpy(_np_i32(a)).to(self._dev())

    # -- gates (device tensors [count, n+1] int32) -------------------------------
    def gate(self, name, ca, cb, out=None, stream=None):
        self._chk(ca, self.words), self._chk(cb, self.words)
        count = ca.shape[0]
        out = self.empty(count) if out is None else out
        self._ck(self.L.tfhe_b200_gate(self.h, GATE_ID[name], out.data_ptr(), ca.data_ptr(), cb.data_ptr(), count,
                                       self._stream(stream)))
        return out

    def gate3(self, gate_id, a, b, c, out=None, stream=None, d=None):
        """A three- (GPC, XOR3, MAJ) or four-input (SUMC, with d) threshold gate on batches of samples."""
        for t in (a, b, c) + ((d,) if d is not None else ()):
            self._chk(t, self.words)
        count = a.shape[0]
        out = self.empty(count) if out is None else out
        op = GateOp(gate_id, count, a.data_ptr(), b.data_ptr(), out.data_ptr(), self.words, self.words, self.words,
                    None, None, None, c.data_ptr(), self.words, None,
                    d.data_ptr() if d is not None else None, self.words, None)
        self._ck(self.L.tfhe_b200_gate_multi(self.h, ctypes.byref(op), 1, self._stream(stream)))
        return out

    def carry_gate(self, g, p, c, out=None, stream=None):
        """out = g | (p & c) in ONE bootstrap (TFHE_B200_GPC; g and p must be mutually exclusive)."""
        for t in (g, p, c):
            self._chk(t, self.words)
        count = g.shape[0]
        out = self.empty(count) if out is None else out
        op = GateOp(GPC, count, g.data_ptr(), p.data_ptr(), out.data_ptr(), self.words, self.words, self.words,
                    None, None, None, c.data_ptr(), self.words, None)
        self._ck(self.L.tfhe_b200_gate_multi(self.h, ctypes.byref(op), 1, self._stream(stream)))
        return out

    def gate2(self, g0, g1, ca, cb, out=None, stream=None):
        count = ca.shape[0]
        out = self.empty(2 * count) if out is None else out
        self._ck(self.L.tfhe_b200_gate2(self.h, GATE_ID[g0], GATE_ID[g1], out.data_ptr(), ca.data_ptr(),
                                        cb.data_ptr(), count, self._stream(stream)))
        return out

    def gate_pair(self, g0, a0, b0, g1, a1, b1, out=None, stream=None):
        count = a0.shape[0]
        out = self.empty(2 * count) if out is None else out
        self._ck(self.L.tfhe_b200_gate_pair(self.h, GATE_ID[g0], a0.data_ptr(), b0.data_ptr(), GATE_ID[g1],
                                            a1.data_ptr(), b1.data_ptr(), out.data_ptr(), count,
                                            self._stream(stream)))
        return out

    def mux(self, a, b, c, out=None, stream=None):
        count = a.shape[0]
        out = self.empty(count) if out is None else out
        self._ck(self.L.tfhe_b200_mux(self.h, out.data_ptr(), a.data_ptr(), b.data_ptr(), c.data_ptr(), count,
                                      self._stream(stream)))
        return out

    def not_(self, ca, out=None, stream=None):
        out = self.empty(ca.shape[0]) if out is None else out
        self._ck(self.L.tfhe_b200_not(self.h, out.data_ptr(), ca.data_ptr(), ca.shape[0], self._stream(stream)))
        return out

    def copy(self, ca, out=None, stream=None):
        out = self.empty(ca.shape[0]) if out is None else out
        self._ck(self.L.tfhe_b200_copy(self.h, out.data_ptr(), ca.data_ptr(), ca.shape[0], self._stream(stream)))
        return out

    def constant(self, value, count, out=None, stream=None):
        out = self.empty(count) if out is None else out
        self._ck(self.L.tfhe_b200_constant(self.h, out.data_ptr(), int(value), count, self._stream(stream)))
        return out

    # -- building blocks ---------------------------------------------------------
    def bootstrap_woks(self, x, mu=MU, stream=None):
        count = x.shape[0]
        u = self.empty(count, self.p.N * self.p.k + 1)
        self._ck(self.L.tfhe_b200_bootstrap_woks(self.h, u.data_ptr(), x.data_ptr(), mu, count, self._stream(stream)))
        return u

    def bootstrap(self, x, mu=MU, stream=None):
        count = x.shape[0]
        out = self.empty(count)
        self._ck(self.L.tfhe_b200_bootstrap(self.h, out.data_ptr(), x.data_ptr(), mu, count, self._stream(stream)))
        return out

    def keyswitch(self, u, stream=None):
        count = u.shape[0]
        out = self.empty(count)
        self._ck(self.L.tfhe_b200_keyswitch(self.h, out.data_ptr(), u.data_ptr(), count, self._stream(stream)))
        return out

    def blind_rotate(self, acc, bara, stream=None):
        """acc [count, k+1, N] (modified in place and returned), bara [count, n_iter]."""
        count, n_iter = bara.shape
        self._ck(self.L.tfhe_b200_blind_rotate(self.h, acc.data_ptr(), bara.data_ptr(), n_iter, count,
                                               self._stream(stream)))
        return acc

    def blind_rotate_and_extract(self, testvect, barb, bara, stream=None):
        count, n_iter = bara.shape
        u = self.empty(count, self.p.N * self.p.k + 1)
        self._ck(self.L.tfhe_b200_blind_rotate_and_extract(self.h, u.data_ptr(), testvect.data_ptr(),
                                                           barb.data_ptr(), bara.data_ptr(), n_iter, count,
                                                           self._stream(stream)))
        return u

    def extern_mul(self, acc, bk_index, stream=None):
        count = acc.shape[0]
        self._ck(self.L.tfhe_b200_extern_mul(self.h, acc.data_ptr(), int(bk_index), count, self._stream(stream)))
        return acc

    # -- host buffers (numpy in, numpy out; copies inside the call) -----------------
    def gate_host(self, name, ca, cb, out=None):
        ca, cb = _np_i32(ca), _np_i32(cb)
        out = np.empty_like(ca) if out is None else out
        self._ck(self.L.tfhe_b200_gate_host(self.h, GATE_ID[name], out.ctypes.data, ca.ctypes.data, cb.ctypes.data,
                                            ca.shape[0]))
        return out

    def mux_host(self, a, b, c):
        a, b, c = _np_i32(a), _np_i32(b), _np_i32(c)
        out = np.empty_like(a)
        self._ck(self.L.tfhe_b200_mux_host(self.h, out.ctypes.data, a.ctypes.data, b.ctypes.data, c.ctypes.data,
                                           a.shape[0]))
        return out


class Circuit:
    """A compiled gate schedule (include/tfhe_b200.h, "Cipher-level circuits")."""

    def __init__(self, engine, kind, *args):
        """engine may be None: the plan can then only be inspected and simulated on plaintext bits."""
        self.eng, self.L = engine, (engine.L if engine is not None else lib())
        h = getattr(self.L, "tfhe_b200_circuit_" + kind)(engine.h if engine is not None else None,
                                                         *[int(a) for a in args])
        if not h:
            raise EngineError("could not build circuit %s%r" % (kind, args))
        self.h = _vp(h)
        self.levels = int(self.L.tfhe_b200_circuit_levels(self.h))
        self.gates = int(self.L.tfhe_b200_circuit_gates(self.h))
        self.level_gates = [int(self.L.tfhe_b200_circuit_level_gates(self.h, i)) for i in range(self.levels)]
        self.out_rows = int(self.L.tfhe_b200_circuit_output_rows(self.h))
        self.operand_rows = [int(self.L.tfhe_b200_circuit_operand_rows(self.h, o))
                             for o in range(int(self.L.tfhe_b200_circuit_operands(self.h)))]

    def run(self, *operands, out=None, stream=None):
        assert len(operands) == len(self.operand_rows)
        for t, rows in zip(operands, self.operand_rows):
            self.eng._chk(t, self.eng.words)
            assert t.numel() == rows * self.eng.words, "operand has the wrong number of samples"
        out = self.eng.empty(self.out_rows) if out is None else out
        ptrs = (_vp * len(operands))(*[t.data_ptr() for t in operands])
        if self.L.tfhe_b200_circuit_run(self.h, out.data_ptr(), ptrs, self.eng._stream(stream)):
            raise EngineError(self.L.tfhe_b200_last_error().decode())
        return out

    def set_graph(self, enable):
        """CUDA-graph replay of the plan's launch sequence (default on; needs a non-default stream)."""
        return bool(self.L.tfhe_b200_circuit_set_graph(self.h, int(bool(enable))))

    @property
    def used_graph(self):
        return bool(self.L.tfhe_b200_circuit_used_graph(self.h))

    @staticmethod
    def run_many(plans, operand_lists, stream=None):
        """K independent plans of one engine, merged level by level into shared launches
        (tfhe_b200_circuit_run_many).  Returns the list of result tensors."""
        eng = plans[0].eng
        outs = [eng.empty(p.out_rows) for p in plans]
        k = len(plans)
        keep = []
        op_arrays = (_vp * k)()
        for i, (p, ops) in enumerate(zip(plans, operand_lists)):
            assert len(ops) == len(p.operand_rows)
            for t, rows in zip(ops, p.operand_rows):
                eng._chk(t, eng.words)
                assert t.numel() == rows * eng.words, "operand has the wrong number of samples"
            arr = (_vp * len(ops))(*[t.data_ptr() for t in ops])
            keep.append(arr)
            op_arrays[i] = ctypes.cast(arr, _vp)
        plan_arr = (_vp * k)(*[p.h for p in plans])
        out_arr = (_vp * k)(*[o.data_ptr() for o in outs])
        if eng.L.tfhe_b200_circuit_run_many(plan_arr, k, out_arr, op_arrays, eng._stream(stream)):
            raise EngineError(eng.L.tfhe_b200_last_error().decode())
        return outs

    def simulate(self, *operand_bits):
        """Plaintext evaluation of the schedule on the host (schedule check; not a compute path)."""
        assert len(operand_bits) == len(self.operand_rows)
        arrs = [np.ascontiguousarray(np.asarray(b, dtype=np.int32).reshape(-1)) for b in operand_bits]
        for a, rows in zip(arrs, self.operand_rows):
            assert a.size == rows
        out = np.zeros(self.out_rows, np.int32)
        ptrs = (_vp * len(arrs))(*[a.ctypes.data for a in arrs])
        if self.L.tfhe_b200_circuit_simulate(self.h, out.ctypes.data, ptrs):
            raise EngineError("circuit simulation failed")
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.tfhe_b200_circuit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiEngine:
    """Several GPUs in one process (tfhe_b200_multi_*, include/tfhe_b200.h): one context per device,
    host batches sharded contiguously, keys uploaded once and copied device to device."""

    def __init__(self, devices=None, params=None):
        self.L = lib()
        self.p = params or default_params()
        h = _vp()
        arr = (ctypes.c_int * len(devices))(*devices) if devices else None
        if self.L.tfhe_b200_multi_create(ctypes.byref(h), ctypes.byref(self.p), arr, len(devices) if devices else 0):
            raise EngineError(self.L.tfhe_b200_multi_last_error().decode())
        self.h = h
        self.ndevices = int(self.L.tfhe_b200_multi_devices(self.h))

    def _ck(self, rc):
        if rc:
            raise EngineError(self.L.tfhe_b200_multi_last_error().decode())

    def load_keys(self, bk_coef, ks):
        bk, k = _np_i32(bk_coef), _np_i32(ks)
        self._ck(self.L.tfhe_b200_multi_load_keys(self.h, bk.ctypes.data, k.ctypes.data))

    def gate_host(self, name, ca, cb, out=None):
        ca, cb = _np_i32(ca), _np_i32(cb)
        out = np.empty_like(ca) if out is None else out
        self._ck(self.L.tfhe_b200_multi_gate_host(self.h, GATE_ID[name], out.ctypes.data, ca.ctypes.data,
                                                  cb.ctypes.data, ca.shape[0]))
        return out

    def mux_host(self, a, b, c):
        a, b, c = _np_i32(a), _np_i32(b), _np_i32(c)
        out = np.empty_like(a)
        self._ck(self.L.tfhe_b200_multi_mux_host(self.h, out.ctypes.data, a.ctypes.data, b.ctypes.data,
                                                 c.ctypes.data, a.shape[0]))
        return out

    @property
    def launch_count(self):
        return int(self.L.tfhe_b200_multi_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.tfhe_b200_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
