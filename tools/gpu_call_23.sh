#!/bin/bash
O=gpurun_out/r2z
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 1200 python bench.py --steps 20 --warmup 3 ) > $O/bench.json 2> $O/bench.err
python -c "import json; d=json.loads(open('$O/bench.json').read()); print('ms/proof', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'marlin', d['marlin']['ms_per_proof'], 'msm', d['msm']['ms_per_msm'], 'setup_s', d['setup_s'], 'launches', d['gpu_launches'])"
tail -3 $O/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 0 ) > $O/bench_ref.json 2> $O/bench_ref.err
python -c "import json; d=json.loads(open('$O/bench_ref.json').read()); print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
