timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/final_bench_n2.json
