#!/bin/bash
# box-side, N GPUs: bench under torchrun with the un-grouped streamed flow (A/B)
N=${1:-2}
mkdir -p gpurun_out
DEMO_QUERY_GROUPS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2f_bench_n${N}_g1.json 2> gpurun_out/r2f_bench_n${N}_g1.err
python - <<PY
import json
j=json.load(open('gpurun_out/r2f_bench_n${N}_g1.json'))
print('groups=1 N=%d value %.0f ms %.2f | e2e %.0f q/s %.2f ms staged %.2f' % (j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['staged_ms_per_step']))
print('e2e stage', {k: round(v,2) for k,v in j['e2e']['stage_ms'].items()})
PY
