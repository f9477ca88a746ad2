#!/bin/bash
# round 2, call 24: pairing tests + Marlin verifier, pairing throughput, NTT unit-twiddle A/B
mkdir -p gpurun_out/r2aa
cd /root/repo
( time timeout 1500 python -m pytest tests/test_gpu_pairing.py tests/test_gpu_marlin_proof.py tests/test_bindings.py -x -q -m gpu ) > gpurun_out/r2aa/pytest_pairing.log 2>&1
tail -15 gpurun_out/r2aa/pytest_pairing.log
timeout 600 python tools/bench_pairing.py > gpurun_out/r2aa/pairing_bench.jsonl 2> gpurun_out/r2aa/pairing_bench.err
cat gpurun_out/r2aa/pairing_bench.jsonl; tail -3 gpurun_out/r2aa/pairing_bench.err
for r in 3 1; do
  ZKB_NTT_RADIX4=$r timeout 600 python tools/bench_ntt.py --steps 5 > gpurun_out/r2aa/ntt_radix4_$r.jsonl 2> gpurun_out/r2aa/ntt_$r.err
  python - <<PY
import json
rows=[json.loads(l) for l in open('gpurun_out/r2aa/ntt_radix4_$r.jsonl') if l.startswith('{')]
for f in ('bls12_381_fr','bn254_fr'):
    print('radix4=$r', f, ' '.join('2^%d:%.3f'%(x['log_n'], x['ms']) for x in rows if x.get('field')==f and x.get('variant')=='fft'))
PY
done
( timeout 600 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu ) > gpurun_out/r2aa/pytest_ntt.log 2>&1; tail -2 gpurun_out/r2aa/pytest_ntt.log
for r in 3 1; do
ZKB_NTT_RADIX4=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2aa/bench_radix4_$r.json 2> gpurun_out/r2aa/err.txt
python -c "import json,sys; d=json.loads(open('gpurun_out/r2aa/bench_radix4_$r.json').read()); print('radix4 $r ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'])"
done
