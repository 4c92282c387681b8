#!/bin/bash
timeout 300 python -m pytest tests/test_ppo.py -q -x -m gpu -k "fused or ppo_improves" 2>&1 | tail -6
timeout 120 python tools/train_ppo.py --envs 16384 --total-steps 6000000 --cuda-graphs 2>&1 | tail -1
