timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
python tools/prof_step.py 148 100 > gpurun_out/r2f_octo_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_octo_kernel -s 1 -c 1 -f -o gpurun_out/r2f_octo_full python tools/prof_step.py 148 100 > gpurun_out/r2f_ncu4.log 2>&1
echo done
