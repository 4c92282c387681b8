#!/bin/bash
# measures the pipeline kernel with two co-resident blocks per SM (slot + ticket role rotation) against the fused kernel
for n in 4096 5120 6144 8192 9472; do
  echo "== n=$n SALP_PIPE_BLOCKS_PER_SM=2"
  SALP_PIPE_BLOCKS_PER_SM=2 DIAG_N=$n timeout 120 python tools/diag_phases.py 2>&1 | tail -2
done
SALP_PIPE_BLOCKS_PER_SM=2 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_long_horizon.py -q -x -m gpu -k "pipeline or handoff or shard" 2>&1 | tail -4
