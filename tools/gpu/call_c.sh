#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/probe_stream.py --large > gpurun_out/r2c_probe.log 2>&1
cat gpurun_out/r2c_probe.log
