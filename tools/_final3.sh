set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 300 python tools/latency_probe.py > gpurun_out/final_latency_probe.log 2>&1
QUAD_NAMES=1 TFHE_B200_LIB=build/variants/pt/libtfhe_b200.so timeout 200 python tools/phase_timing.py 1 148 > gpurun_out/final_phase_timing.log 2>&1
python tools/prof_step.py 148 100 > gpurun_out/r2f_octo_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_octo_kernel -s 1 -c 1 -f -o gpurun_out/r2f_octo_full python tools/prof_step.py 148 100 > gpurun_out/r2f_ncu4.log 2>&1
tail -3 gpurun_out/final_pytest.log; tail -2 gpurun_out/final_smoke.log; cat gpurun_out/final_latency_probe.log | tail -8
