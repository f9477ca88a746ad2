#!/bin/bash
# round 2, call 25: short-MSM batch path, cyclotomic squarings, verify sub-record, sanitizer over the new kernels
O=gpurun_out/r2ab
mkdir -p $O
cd /root/repo
( time timeout 1500 python -m pytest tests/test_gpu_pairing.py tests/test_gpu_msm.py tests/test_gpu_marlin_proof.py tests/test_gpu_kzg10.py -x -q -m gpu ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
timeout 600 python tools/bench_pairing.py > $O/pairing_bench.jsonl 2> $O/pairing_bench.err
cat $O/pairing_bench.jsonl; tail -3 $O/pairing_bench.err
TOOLS="memcheck" TMO=1200 bash tools/sanitize.sh $O/sanitizer > $O/sanitize.out 2>&1; tail -4 $O/sanitize.out
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('ms/proof',d['ms_per_step'],'e2e',d['e2e']['value']);print(json.dumps(d.get('verify')))"
tail -3 $O/bench.err
