#!/bin/bash
mkdir -p gpurun_out
for g in 0 100000; do DEMO_CM_SMALL_G=$g timeout 300 python tools/time_count_small.py 2>&1 | tail -8; done | tee gpurun_out/r2aq_small.log
