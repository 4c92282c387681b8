#!/bin/bash
# round-2 GPU session 1: parity suite on the new kernels + small-batch timing A/B
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_gpu.txt
python -m pytest tests -m gpu -q -x --timeout 1500 > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
SALP_PIPE_VARIANT=3 python tools/diag_small_batch.py > gpurun_out/s1_diag_v3.log 2>&1
python tools/diag_small_batch.py > gpurun_out/s1_diag_v4.log 2>&1
tail -30 gpurun_out/s1_diag_v3.log gpurun_out/s1_diag_v4.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; tail -c 1500 gpurun_out/s1_bench.json
