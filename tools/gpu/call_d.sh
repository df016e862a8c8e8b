#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2d_overlap.log; : > $out
for cfg in "0 0" "72 16" "72 32" "70 32" "70 64" "66 64" "72 8"; do
  set -- $cfg
  echo "=== DEMO_PAIRS=$1 DEMO_STREAM_TOTAL=$2" >> $out
  DEMO_PAIRS=$1 DEMO_STREAM_TOTAL=$2 timeout 200 python tools/probe_stream.py --large --overlap-only 2>&1 | grep -v "^problem\|^plan" >> $out
done
cat $out
