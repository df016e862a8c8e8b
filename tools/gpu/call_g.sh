#!/bin/bash
# DRAM traffic of the count GEMM at 20k x 1M for schedule variants (ncu metrics-only)
mkdir -p gpurun_out
out=gpurun_out/r2g_traffic.log; : > $out
for cfg in "DEMO_ADJ=1" "DEMO_ADJ=0" "DEMO_ADJ=0 DEMO_PACE=0" "DEMO_ADJ=1 DEMO_PACE=0" "DEMO_ADJ=1 DEMO_CHUNK_TILES=4" "DEMO_ADJ=1 DEMO_GROUP_M=74"; do
  echo "=== $cfg" >> $out
  env $cfg timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sqdist_gemm2_kernel -s 1 -c 1 \
    python tools/exp_count.py 20000 1000000 1536 1 2>&1 | grep -E "dram__bytes_read|gpu__time_duration|hit_rate|count:" >> $out
done
cat $out
