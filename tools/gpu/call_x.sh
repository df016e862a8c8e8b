#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/rerank_stats.py > gpurun_out/r2x_stats.log 2>&1; cat gpurun_out/r2x_stats.log | tail -4
