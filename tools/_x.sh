for v in base mac4 base mac4; do
  if [ "$v" = base ]; then lib=cpu-gpu-tfhe_b200/libtfhe_b200.so; else lib=build/variants/$v/libtfhe_b200.so; fi
  echo "== $v"; TFHE_B200_LIB=$lib timeout 300 python tools/latency_probe.py 2>&1 | grep "count    1:\|count  148:"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -2
TFHE_B200_LIB=build/variants/mac4/libtfhe_b200.so timeout 600 python -m pytest tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -2
