timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/final_bench_n8.json 2> gpurun_out/final_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['strong']['value'], d['strong']['ms_per_step'], d['latency']['matmul16x16_8bit'], d['latency']['matmul16x16_8bit_reference_schedule']['ms'])
PY
