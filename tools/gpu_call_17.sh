#!/bin/bash
# device-resident buffers used in place, no sync when the output stays on the device: parity + Marlin / PLONK timings
O=gpurun_out/r2s
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 600 python tools/prof_marlin.py > $O/prof_marlin.txt 2>&1; head -18 $O/prof_marlin.txt
( time timeout 1200 python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
python -c "import json; d=json.loads(open('$O/bench.json').read()); print('ms/proof', d['ms_per_step'], 'e2e', d['e2e']['value'], 'marlin', d['marlin']['ms_per_proof'], 'msm', d['msm']['ms_per_msm'])"
timeout 600 python tools/bench_plonk.py --log-n 16 > $O/plonk16.json 2> $O/plonk16.err; cat $O/plonk16.json
timeout 600 python tools/bench_plonk.py --log-n 18 > $O/plonk18.json 2> $O/plonk18.err; cat $O/plonk18.json
timeout 600 python tools/bench_plonk.py --log-n 20 --steps 2 > $O/plonk20.json 2> $O/plonk20.err; cat $O/plonk20.json; tail -2 $O/plonk20.err
