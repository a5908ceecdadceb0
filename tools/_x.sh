timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py tests/test_gpu_circuits.py tests/test_gpu_compat.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/latency_probe.py 2>&1 | tail -8
