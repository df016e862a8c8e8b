#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2e_e2e.log; : > $out
for cfg in "8 64" "4 64" "12 96" "0 64"; do
  set -- $cfg
  echo "=== DEMO_RESERVE_SMS=$1 DEMO_STREAM_TOTAL=$2" >> $out
  DEMO_RESERVE_SMS=$1 DEMO_STREAM_TOTAL=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other 2>>$out | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('value %.0f ms %.2f | e2e %.0f q/s %.2f ms staged %.2f identical %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['staged_ms_per_step'], j['e2e']['identical_to_device_resident_result']))
print({k: round(v,2) for k,v in j['e2e']['stage_ms'].items()})" >> $out
done
cat $out
