#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_long_horizon.py -q -x -m gpu -k "pipeline or handoff" 2>&1 | tail -4
timeout 100 python tools/diag_phases.py
timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-sweep | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
