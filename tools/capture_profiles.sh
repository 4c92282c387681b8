#!/bin/bash
# Round-end profile capture (one GPU).  Each ncu pass runs only after the same command exited 0
# without ncu.  Outputs go to gpurun_out/ (merged back by gpurun); summaries are made afterwards
# with tools/ncu_summary.py and committed under profiles/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-sweep --no-cpu-baseline --no-e2e"
for n in 4096 262144; do
  $B --envs $n > gpurun_out/plain_$n.json 2> gpurun_out/plain_$n.err || { echo "plain run failed for $n"; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches_$n.csv $B --envs $n > gpurun_out/ncu_l_$n.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:salp_step_kernel -s 6 -c 1 \
      -o gpurun_out/step_$n -f $B --envs $n > gpurun_out/ncu_f_$n.log 2>&1
done
ls -la gpurun_out
