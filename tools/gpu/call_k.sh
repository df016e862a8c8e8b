#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2k_pytest.log
cat gpurun_out/r2k_pytest.log
out=gpurun_out/r2k_e2e.log; : > $out
for g in 4 1 8 6; do
  echo "=== DEMO_QUERY_GROUPS=$g" >> $out
  DEMO_QUERY_GROUPS=$g timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other 2>>$out | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('value %.0f ms %.2f | e2e %.0f q/s %.2f ms staged %.2f identical %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['staged_ms_per_step'], j['e2e']['identical_to_device_resident_result']))
print({k: round(v,2) for k,v in j['e2e']['stage_ms'].items()})" >> $out
done
cat $out
