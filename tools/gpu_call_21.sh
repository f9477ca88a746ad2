#!/bin/bash
# fused two-axis reduction passes + mid-size window rule: parity, stand-alone MSM timings A/B, proof, Marlin
O=gpurun_out/r2w
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for f in 0 1; do
  echo "fused=$f"
  ZKB_REDUCE_FUSED=$f timeout 300 python tools/exp_pair.py --levels 0 --batch 0 --steps 5 2>&1 | grep -o '"group": [12], "log_n": [0-9]*\|"ms_median": [0-9.]*' | paste - - | sed 's/"//g'
  ZKB_REDUCE_FUSED=$f timeout 300 python tools/exp_pair.py --log-n 17 --levels 0 --batch 0 --steps 5 2>&1 | grep -o '"group": [12], "log_n": [0-9]*\|"ms_median": [0-9.]*' | paste - - | sed 's/"//g'
  ZKB_REDUCE_FUSED=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_fused$f.json 2> $O/err.txt
  python -c "import json,sys; d=json.loads(open('$O/bench_fused$f.json').read()); print('fused $f ms/proof', d['ms_per_step'], 'e2e', d['e2e']['value'], 'verified', d['verified_in_exponent'])"
done
( time timeout 1200 python bench.py --steps 20 --warmup 3 ) > $O/bench.json 2> $O/bench.err
python -c "import json; d=json.loads(open('$O/bench.json').read()); print('ms/proof', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'marlin', d['marlin']['ms_per_proof'], 'msm', d['msm']['ms_per_msm'])"
