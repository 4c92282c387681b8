"""GPU: RecurrentPPO update time per iteration (8192 envs, T = 32, 10 epochs x 16 minibatches, CUDA graphs)
with the LSTM sequence function (lstm_seq.py) and with nn.LSTMCell stepped T times.
   python tools/diag_rppo_update.py [N]"""
import json
import sys

import torch

from grasp_lab_salp_b200 import SalpBatch, default_params
from grasp_lab_salp_b200.ppo import DeviceEnv, PPOConfig, RecurrentPPO

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for fused in (True, False):
    batch = SalpBatch(n, default_params(), seed=0)
    algo = RecurrentPPO(DeviceEnv(batch), PPOConfig(n_steps=32, batch_size=16384, cuda_graphs=True, seed=0, fused_sequence=fused))
    rows = []
    algo.learn(4 * 32 * n, log=rows.append)
    torch.cuda.synchronize()
    batch.check()
    r = rows[-1]
    print(json.dumps({"envs": n, "lstm_sequence_function": fused, "update_seconds": r["update_seconds"],
                      "rollout_seconds": r["rollout_seconds"], "approx_kl": r["approx_kl"], "value_loss": r["value_loss"],
                      "update_seconds_all": [x["update_seconds"] for x in rows]}), flush=True)
