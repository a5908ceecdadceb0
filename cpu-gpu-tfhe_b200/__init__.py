"""B200-native TFHE gate-bootstrapping engine — Python host layer.

This package only *binds* the C ABI of ``libtfhe_b200.so`` (include/tfhe_b200.h,
include/tfhe_compat.h) with ctypes and uses PyTorch for device memory and
streams.  All arithmetic happens in the hand-written sm_100a kernels under
``csrc/``; there is no Python or CPU fallback — if the library or a CUDA device
is missing the calls raise.

The directory name contains a hyphen, so import it through
``__graft_entry__.load_package()`` (registers it as ``cpu_gpu_tfhe_b200``).
"""
from . import binding  # noqa: F401
from .binding import (  # noqa: F401
    CMP,
    GATES,
    SHIFT,
    GATE_ID,
    Circuit,
    Engine,
    MultiEngine,
    EngineError,
    Params,
    SecretKeys,
    decrypt_bits,
    encrypt_bits,
    keygen,
    phases,
    default_params,
    device_count,
    lib,
    lib_path,
    measure_fp64_peak,
    read_ciphertexts,
    read_key_file,
    write_ciphertexts,
    write_cloud_key,
    write_secret_key,
)
