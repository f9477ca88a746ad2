#!/bin/bash
# full GPU suite on the round-2 tree (batched MSM entry, PC::commit batching, 4-way reduction passes), A/B of the G2
# accumulation with the multiplication as a call, bench line
O=gpurun_out/r2l
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
for rep in 1 2; do for call in 0 2; do
  ZKB_ACC_CALL=$call timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_call${call}_$rep.json 2> $O/err.txt
  python -c "import json,sys; d=json.loads(open('$O/bench_call${call}_$rep.json').read()); print('acc_call $call rep $rep ms/proof', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done; done
( time timeout 1200 python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
tail -c 1200 $O/bench.json; tail -5 $O/bench.err
