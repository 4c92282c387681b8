#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/train_ppo.py --envs 16384 --total-steps 40000000 --cuda-graphs --out gpurun_out/r02_ppo_mlp_16384envs.jsonl > gpurun_out/train_mlp.log 2>&1; tail -1 gpurun_out/train_mlp.log
timeout 900 python tools/train_ppo.py --recurrent --envs 8192 --total-steps 40000000 --batch-size 16384 --cuda-graphs --out gpurun_out/r02_rppo_lstm_8192envs.jsonl > gpurun_out/train_rppo.log 2>&1; tail -1 gpurun_out/train_rppo.log
python - <<'PY'
import json
for f in ('gpurun_out/r02_ppo_mlp_16384envs.jsonl','gpurun_out/r02_rppo_lstm_8192envs.jsonl'):
    rows=[json.loads(l) for l in open(f) if l.strip().startswith('{')]
    print(f, len(rows))
    for r in rows[::max(1,len(rows)//10)]+rows[-1:]:
        print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('env_steps','mean_episode_return','mean_episode_length','success_rate','approx_kl','clip_fraction','value_loss','wall_seconds')})
PY
