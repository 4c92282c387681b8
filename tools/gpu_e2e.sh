#!/bin/bash
# e2e (host buffers) against device-timed throughput at the large sweep sizes, tapered vs equal host ranges
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "chunked_host" 2>&1 | tail -2
for n in 262144 1048576; do
  for t in 1 0; do
    SALP_HOST_RANGE_TAIL=$t timeout 300 python bench.py --envs $n --steps 30 --warmup 5 --no-cpu-baseline --no-sweep | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n=$n taper=$t device', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'ratio', round(d['e2e']['value']/d['value'],3))"
  done
done
