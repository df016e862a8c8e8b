#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2b.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/launches_r2b.log 2>&1
tail -2 gpurun_out/launches_r2b.log | cut -c1-300
