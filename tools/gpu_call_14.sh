#!/bin/bash
# NTT with radix-4 steps (two stages per shared-memory round trip): parity, sweep A/B, ncu of the pass kernel
O=gpurun_out/r2p
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_groth16.py tests/test_gpu_msm.py -x -q -m gpu ) > $O/pytest_ntt.log 2>&1
tail -4 $O/pytest_ntt.log
for r in 0 1; do
  ZKB_NTT_RADIX4=$r timeout 600 python tools/bench_ntt.py --steps 5 > $O/ntt_radix4_$r.jsonl 2> $O/ntt_$r.err
  python - <<PY
import json
rows=[json.loads(l) for l in open('$O/ntt_radix4_$r.jsonl') if l.startswith('{')]
for f in ('bls12_381_fr','bn254_fr'):
    print('radix4=$r', f, ' '.join('2^%d:%.3f'%(x['log_n'], x['ms']) for x in rows if x.get('field')==f and x.get('variant')=='fft'))
PY
done
for r in 0 1; do
ZKB_NTT_RADIX4=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_radix4_$r.json 2> $O/err.txt
python -c "import json,sys; d=json.loads(open('$O/bench_radix4_$r.json').read()); print('radix4 $r ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'])"
done
