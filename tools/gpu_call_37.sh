#!/bin/bash
# round 2, call 37: final 1-GPU bench line (all sub-records) on the final tree + smoke
O=gpurun_out/r2ak
mkdir -p $O
cd /root/repo
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('ms/proof',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'identical',d.get('gpu_proof_identical_to_cpu_port'));m=d['marlin'];print('marlin',m['ms_per_proof'],m['verified_on_gpu']);print([ (r['field'],r['log_n'],round(r['fft']['ms'],3)) for r in d['ntt']['sizes']]);print('msm',d['msm']['ms_per_msm']);print([(r['curve'],round(r['checks_per_s']),r['products_are_one']) for r in d['verify']['runs']]);print(d['roofline']['frac'], d.get('roofline_integer',{}).get('frac'))"
tail -2 $O/bench.err
