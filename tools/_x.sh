timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'pageable',d['e2e']['pageable']['value'])"
python - <<"PY"
import torch,time
x=torch.empty(256*1024*1024,dtype=torch.uint8).pin_memory(); d=torch.empty_like(x,device='cuda')
for _ in range(2): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); t=time.time(); d.copy_(x,non_blocking=True); torch.cuda.synchronize(); print('H2D GB/s', 0.268/(time.time()-t))
t=time.time(); x.copy_(d,non_blocking=True); torch.cuda.synchronize(); print('D2H GB/s', 0.268/(time.time()-t))
PY
