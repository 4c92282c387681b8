#!/bin/bash
# Round-end capture (one GPU): full GPU test suite, the default bench line of both arms, launch lists
# and one `ncu --set full` capture of the step kernel at 4096 and 262144 envs (each ncu pass only
# after the same command exited 0 without ncu), the micro-benchmarks, the config-2 trajectory run.
# Outputs go to gpurun_out/ (merged back by gpurun); summaries are made afterwards with
# tools/ncu_summary.py and committed under profiles/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c_pytest.log; tail -3 gpurun_out/c_pytest.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c_bench_reference.json 2> gpurun_out/c_bench_reference.err; echo "reference arm rc $?"
timeout 900 python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc $?"
tools/_bin/ubench_fp32 > gpurun_out/c_ubench_fp32.txt 2>&1
tools/_bin/ubench_substep > gpurun_out/c_ubench_substep.txt 2>&1
B="python bench.py --steps 8 --warmup 3 --no-sweep --no-cpu-baseline --no-e2e"
for n in 4096 262144; do
  timeout 200 $B --envs $n > gpurun_out/c_plain_$n.json 2> gpurun_out/c_plain_$n.err || { echo "plain run failed for $n"; continue; }
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/c_launches_$n.csv $B --envs $n > gpurun_out/c_ncu_l_$n.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:salp_step_kernel -s 6 -c 1 \
      -o gpurun_out/c_step_$n -f $B --envs $n > gpurun_out/c_ncu_f_$n.log 2>&1
done
timeout 600 python tools/compare_trajectories.py --envs 4096 --steps 10000 --out gpurun_out/c_traj_equiv.json > gpurun_out/c_traj.log 2>&1; tail -2 gpurun_out/c_traj.log
ls -la gpurun_out | grep " c_"
