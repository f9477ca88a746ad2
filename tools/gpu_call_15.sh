#!/bin/bash
# NTT radix-4 x occupancy target sweep; PLONK resident tests + bench; curve tests
O=gpurun_out/r2q
mkdir -p $O
for occ in 2 3 4; do for r in 0 1; do
  ZKB_NTT_OCC=$occ ZKB_NTT_RADIX4=$r timeout 600 python tools/bench_ntt.py --steps 5 --min-log 20 --max-log 24 > $O/ntt_occ${occ}_r$r.jsonl 2> $O/ntt.err
  python - <<PY
import json
rows=[json.loads(l) for l in open('$O/ntt_occ${occ}_r$r.jsonl') if l.startswith('{')]
for f in ('bls12_381_fr','bn254_fr'):
    print('occ=$occ radix4=$r', f, ' '.join('2^%d:%.3f'%(x['log_n'], x['ms']) for x in rows if x.get('field')==f and x.get('variant')=='fft'))
PY
done; done
( time timeout 900 python -m pytest tests/test_gpu_plonk.py tests/test_gpu_curve.py tests/test_gpu_ntt.py -x -q -m gpu ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
timeout 600 python tools/bench_plonk.py --log-n 16 > $O/plonk16.json 2> $O/plonk16.err; cat $O/plonk16.json; tail -3 $O/plonk16.err
timeout 600 python tools/bench_plonk.py --log-n 18 > $O/plonk18.json 2> $O/plonk18.err; cat $O/plonk18.json; tail -3 $O/plonk18.err
timeout 600 python tools/bench_plonk.py --log-n 16 --host-buffers > $O/plonk16_host.json 2> $O/plonk16_host.err; cat $O/plonk16_host.json
