#!/bin/bash
# Round-2, second-session capture (one GPU, ~9 min): full GPU test suite, both bench arms, the pipeline
# kernel's clock stamps and phase timings, launch list + one `ncu --set full` capture of the step kernel at
# 4096 envs and of the LSTM cell at 8192 envs (each ncu pass only after the same command exited 0 without
# ncu).  Outputs: gpurun_out/d_*; summaries are made afterwards with tools/ncu_summary.py -> profiles/.
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/d_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/d_pytest.log; tail -3 gpurun_out/d_pytest.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/d_bench_reference.json 2> gpurun_out/d_bench_reference.err; echo "reference arm rc $?"
timeout 900 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc $?"
timeout 100 python tools/diag_stamps.py > gpurun_out/d_stamps.txt 2>&1
for n in 4096 8192; do DIAG_N=$n timeout 100 python tools/diag_phases.py >> gpurun_out/d_phases.txt 2>&1; done
timeout 100 python tools/diag_lstm.py 100 8192 65536 > gpurun_out/d_lstm.jsonl 2>&1
timeout 100 python tools/diag_rppo_rollout.py 8192 > gpurun_out/d_rppo_rollout.jsonl 2>&1
B="python bench.py --steps 8 --warmup 3 --no-sweep --no-cpu-baseline --no-e2e"
n=4096
timeout 200 $B --envs $n > gpurun_out/d_plain_$n.json 2> gpurun_out/d_plain_$n.err || echo "plain run failed for $n"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/d_launches_$n.csv $B --envs $n > gpurun_out/d_ncu_l_$n.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:salp_step_kernel -s 6 -c 1 \
    -o gpurun_out/d_step_$n -f $B --envs $n > gpurun_out/d_ncu_f_$n.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:salp_lstm_cell_kernel -s 3 -c 1 \
    -o gpurun_out/d_lstm_cell_8192 -f python tools/diag_lstm.py 8192 > gpurun_out/d_ncu_f_lstm.log 2>&1
ls -la gpurun_out | grep " d_"
