#!/bin/bash
# round 2, call 26: batched evaluations / openings in the Marlin prover
O=gpurun_out/r2ac
mkdir -p $O
cd /root/repo
( time timeout 1500 python -m pytest tests/test_gpu_kzg10.py tests/test_gpu_marlin_proof.py tests/test_gpu_marlin.py tests/test_gpu_plonk.py tests/test_gpu_sharded.py -x -q -m gpu ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
timeout 600 python tools/prof_marlin.py > $O/prof_marlin.txt 2>&1; head -22 $O/prof_marlin.txt
TOOLS="memcheck" TMO=1200 bash tools/sanitize.sh $O/sanitizer > $O/sanitize.out 2>&1; tail -4 $O/sanitize.out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('ms/proof',d['ms_per_step']);m=d['marlin'];print('marlin',m['ms_per_proof'],m['verified_on_gpu'],m['gpu_launches_per_rank'])"
tail -3 $O/bench.err
