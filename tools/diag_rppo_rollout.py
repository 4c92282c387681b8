"""GPU: RecurrentPPO rollout time (T = 32 steps, CUDA graph) with the tcgen05 LSTM cells and with the
fp32 torch cells.   python tools/diag_rppo_rollout.py [N]"""
import json
import sys
import time

import torch

from grasp_lab_salp_b200 import SalpBatch, default_params
from grasp_lab_salp_b200.ppo import DeviceEnv, PPOConfig, RecurrentPPO

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for fused in (True, False):
    batch = SalpBatch(n, default_params(), seed=0)
    algo = RecurrentPPO(DeviceEnv(batch), PPOConfig(n_steps=32, batch_size=16384, cuda_graphs=True, seed=0, fused_policy=fused))
    for _ in range(3):
        algo.collect()                    # eager, capture, first replay
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        algo.collect()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    batch.check()
    print(json.dumps({"envs": n, "tensor_core_lstm": fused, "rollout_ms_per_32_steps": ms,
                      "env_steps_per_sec_rollout_only": n * 32 / ms * 1e3}), flush=True)
