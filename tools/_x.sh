timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q -k "host or multi" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'pageable',d['e2e']['pageable']['value'],d['e2e']['pageable']['ms_per_step'])"
