#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_timeline.py > gpurun_out/r2af_timeline.log 2>&1
tail -3 gpurun_out/r2af_timeline.log
