#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_eval_auto.py > gpurun_out/r2m_diag.log 2>&1
cat gpurun_out/r2m_diag.log
