#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_eval_auto.py > gpurun_out/r2m_diag.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python tools/diag_eval_auto.py 2>&1 | grep -E "^  [a-zA-Z<]|gpu__time" | paste - - | awk '{print $1, $(NF-1), $NF}' | sort | uniq -c | sort -k4 -n | tail -30 >> gpurun_out/r2m_diag.log
cat gpurun_out/r2m_diag.log
