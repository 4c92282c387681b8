#!/bin/bash
# ncu evidence for the tensor-core LSTM cell: plain run first, then the launch list and one --set full
# capture of salp_lstm_cell_kernel per batch size (outputs in gpurun_out/, summaries go to profiles/)
export PYTHONPATH=.
for n in ${LSTM_SIZES:-8192 65536}; do
  timeout 100 python tools/diag_lstm.py $n > gpurun_out/l_plain_$n.json 2> gpurun_out/l_plain_$n.err || { echo "plain run failed for $n"; continue; }
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:salp_lstm -c 40 --csv \
      --log-file gpurun_out/l_launches_$n.csv python tools/diag_lstm.py $n > gpurun_out/l_ncu_l_$n.log 2>&1
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:salp_lstm_cell_kernel -s 3 -c 1 \
      -o gpurun_out/l_cell_$n -f python tools/diag_lstm.py $n > gpurun_out/l_ncu_f_$n.log 2>&1
done
ls -la gpurun_out | grep " l_"
