#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2o_pytest.log
cat gpurun_out/r2o_pytest.log
for g in 4 1; do
DEMO_QUERY_GROUPS=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2o_bench_n2_g$g.json 2> gpurun_out/r2o_bench_n2_g$g.err
tail -3 gpurun_out/r2o_bench_n2_g$g.err
python - <<PY
import json
j=json.load(open('gpurun_out/r2o_bench_n2_g$g.json'))
print('groups $g: N=%d value %.0f ms %.2f | e2e %.0f q/s %.2f ms staged %.2f same %s' % (j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['staged_ms_per_step'], j['e2e']['identical_to_device_resident_result']))
print('e2e stage', {k: round(v,2) for k,v in j['e2e']['stage_ms'].items()})
print(j.get('multi_gpu_checks'))
PY
done
