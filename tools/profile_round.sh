#!/bin/bash
# Produces the ncu evidence committed under profiles/ (run on the GPU box through gpurun):
#   launches.csv  per-launch device time of `bench.py --steps 1 --warmup 3` (cold cache, serialised)
#   traffic.csv   DRAM / L2 bytes of one 65536-gate launch of each hot kernel
#   br_full.ncu-rep  --set full capture of the blind-rotation kernel (batch 4736 = 8 waves)
set -x
tag=${1:-r1s3}
out=gpurun_out
cmd="python bench.py --steps 1 --warmup 3 --no-latency --no-cpu-baseline"
$cmd > $out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'blind_rotate_kernel|keyswitch_mma_kernel' -s 6 -c 2 --csv --log-file $out/${tag}_traffic.csv $cmd > $out/${tag}_ncu2.log 2>&1
python tools/quick_bench.py 4736 > $out/${tag}_qb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 2 -c 1 -f -o $out/${tag}_br_full python tools/quick_bench.py 4736 > $out/${tag}_ncu3.log 2>&1
ls -la $out
