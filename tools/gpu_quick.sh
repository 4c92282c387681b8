#!/bin/bash
# quick check of the pipeline kernel: its parity tests, the phase timings, a short bench
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_long_horizon.py -q -x -m gpu -k "pipeline or handoff or shard" 2>&1 | tail -4
for n in ${DIAG_SIZES:-4096}; do
  DIAG_N=$n timeout 120 python tools/diag_phases.py 2>&1 | tail -2
done
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-sweep | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
