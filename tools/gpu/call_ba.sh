#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2ba_bench.json 2> gpurun_out/r2ba_bench.err; echo bench rc=$?
