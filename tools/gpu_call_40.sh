#!/bin/bash
# round 2, call 40: NTT pass kernel back to the r2q code (the unit-twiddle skip cost registers: +7 % per transform in the
# bench sub-record); whole GPU suite, NTT sweep, final bench line
O=gpurun_out/r2an
mkdir -p $O
cd /root/repo
timeout 300 python tools/bench_ntt.py --steps 5 > $O/ntt.jsonl 2> $O/ntt.err
python - <<PY
import json
rows=[json.loads(l) for l in open('$O/ntt.jsonl') if l.startswith('{')]
for f in ('bls12_381_fr','bn254_fr'):
    print(f, ' '.join('2^%d:%.3f'%(x['log_n'], x['ms']) for x in rows if x.get('field')==f and x.get('variant')=='fft'))
PY
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print('ms/proof',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'identical',d.get('gpu_proof_identical_to_cpu_port'));m=d['marlin'];print('marlin',m['ms_per_proof'],m['verified_on_gpu']);print([ (r['field'],r['log_n'],round(r['fft']['ms'],3)) for r in d['ntt']['sizes']]);print('msm',d['msm']['ms_per_msm']);print([(r['curve'],round(r['checks_per_s']),r['products_are_one']) for r in d['verify']['runs']])"
tail -2 $O/bench.err
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
