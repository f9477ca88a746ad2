#!/bin/bash
# round 2, call 34: multi-GPU bench line of the final tree (sharded proof / MSM / Marlin sub-records)
N=$1
O=gpurun_out/r2ai_${N}gpu
mkdir -p $O
cd /root/repo
( time NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('N=$N value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'sharded_proof ms', round(d['sharded_proof']['ms_per_proof'],2), 'msm ms', round(d['msm']['ms_per_msm'],2), 'marlin ms', round(d['marlin']['ms_per_proof'],1), d['marlin'].get('verified_on_gpu'))
PY
tail -3 $O/bench.err
