#!/bin/bash
# quick GPU check: pipeline-kernel parity tests + small-batch timings + bench line
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_env_surfaces.py -q -x -k "chunked or zero_copy" 2>&1 | tail -5
#timeout 100 python tools/diag_phases.py
timeout 400 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], )"
python -c "
import json; d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
for r in d['sweep']: print(r['envs_per_gpu'], '%.1f M'%(r['env_steps_per_sec']/1e6), 'e2e %.1f M'%(r.get('e2e_env_steps_per_sec',0)/1e6), 'ratio %.2f'%(r.get('e2e_env_steps_per_sec',0)/r['env_steps_per_sec']))"
