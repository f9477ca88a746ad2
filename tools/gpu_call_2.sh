#!/bin/bash
# round-2 GPU call B: parity suite again (pair levels fixed, Marlin driver), pair-level sweep with full-width scalars,
# whole proof with pair levels, ncu capture of the level / accumulate kernels, all four sanitizer tools on the full target.
O=gpurun_out/r2b
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 600 python tools/exp_pair.py > $O/exp_pair.jsonl 2>&1
for lv in 0 1 2 3; do
  ZKB_MSM_PAIR_LEVELS=$lv timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_levels$lv.json 2> $O/bench_levels$lv.err
  python -c "import json,sys; d=json.loads(open('$O/bench_levels$lv.json').read()); print('levels $lv ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pair_level|k_accumulate' -s 8 -c 6 -o $O/pair_ncu \
   python tools/exp_pair.py --levels 2 --scales 0 --groups 1 --steps 1 > $O/ncu_pair.log 2>&1
tail -3 $O/ncu_pair.log
TMO=600 bash tools/sanitize.sh $O/sanitizer > $O/sanitize_summary.txt 2>&1
cat $O/sanitize_summary.txt
