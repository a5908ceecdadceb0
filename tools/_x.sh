timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/latency_probe.py 2>&1 | tail -8
QUAD_NAMES=1 TFHE_B200_LIB=build/variants/pt/libtfhe_b200.so timeout 200 python tools/phase_timing.py 1 148
