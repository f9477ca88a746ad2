#!/bin/bash
# round 2, call 29: whole GPU suite, smoke, final 1-GPU bench line and reference arm on the final tree
O=gpurun_out/r2af
mkdir -p $O
cd /root/repo
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('ms/proof',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value']);m=d['marlin'];print('marlin',m['ms_per_proof'],m['verified_on_gpu']);print([ (r['field'],r['log_n'],round(r['fft']['ms'],3)) for r in d['ntt']['sizes']]);print('msm',d['msm']['ms_per_msm']);print([(r['curve'],round(r['checks_per_s'])) for r in d['verify']['runs']])"
tail -2 $O/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-400 $O/bench_ref.json
