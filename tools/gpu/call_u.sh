#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sqdist_gemm|count_matrix|resolve|prep_rows|build_thr|block_flags" -c 400 --csv --log-file gpurun_out/launches_r171_r2.csv python tools/profile_r171.py > gpurun_out/launches_r171_r2.log 2>&1
tail -2 gpurun_out/launches_r171_r2.log | cut -c1-200
