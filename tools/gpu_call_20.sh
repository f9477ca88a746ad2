#!/bin/bash
# after the mid-size window rule: multi-GPU bench line (sharded proof / MSM / Marlin sub-records), with and without
# low-priority accumulation streams for the sharded proof
N=$1
O=gpurun_out/r2v_${N}gpu
mkdir -p $O
for bulk in 0 1; do
( time ZKB_BULK=$bulk NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 ) > $O/bench_bulk$bulk.json 2> $O/bench_bulk$bulk.err
python - <<PY
import json
d=json.loads(open('$O/bench_bulk$bulk.json').read())
print('bulk=$bulk N=$N value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'sharded_proof ms', round(d['sharded_proof']['ms_per_proof'],2), 'msm ms', round(d['msm']['ms_per_msm'],2), 'marlin ms', round(d['marlin']['ms_per_proof'],1))
PY
done
