#!/bin/bash
# sorts of the four assignment MSMs enqueued before any accumulation (run_split): parity + A/B + timeline
O=gpurun_out/r2y
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_groth16.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py -x -q -m gpu ) > $O/pytest.log 2>&1
tail -3 $O/pytest.log
for rep in 1 2; do for sf in 0 1; do
  ZKB_SORTS_FIRST=$sf timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_sf${sf}_$rep.json 2> $O/err.txt
  python -c "import json,sys; d=json.loads(open('$O/bench_sf${sf}_$rep.json').read()); print('sorts_first $sf rep $rep ms/proof', d['ms_per_step'], 'e2e', d['e2e']['value'], d['verified_in_exponent'])"
done; done
timeout 300 python tools/timeline.py --out $O/timeline_sorts_first.txt > $O/tl.log 2>&1
head -3 $O/timeline_sorts_first.txt
