#!/bin/bash
# batched-affine accumulation v3 (two levels of chains, operand look-ahead): parity, timings, sweep, reduction with 4-way passes
O=gpurun_out/r2j
mkdir -p $O
( time ZKB_MSM_BATCH=1 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu ) > $O/pytest_msm.log 2>&1
tail -5 $O/pytest_msm.log
timeout 300 python tools/exp_pair.py --levels 0 --batch 0 1 --steps 5 > $O/exp_batch.jsonl 2> $O/exp_batch.err
cat $O/exp_batch.jsonl; tail -3 $O/exp_batch.err
for cfg in "4 8 8" "4 6 6" "4 12 4" "4 16 4" "5 8 8" "3 8 8" "4 4 8"; do
  set -- $cfg
  echo "bps $1 l0 $2 l1 $3"
  ZKB_BATCH_BPS=$1 ZKB_BATCH_L0=$2 ZKB_BATCH_L1=$3 timeout 300 python tools/exp_pair.py --levels 0 --batch 1 --groups 1 --steps 3 2>&1 | tail -1
done
for k in 2 4 8; do
  echo "seg_k $k"; ZKB_SEG_K=$k timeout 300 python tools/exp_pair.py --levels 0 --batch 0 --groups 1 2 --steps 3 2>&1 | tail -2
done
ZKB_MSM_BATCH=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_batch1.json 2> $O/bench_batch1.err
python -c "import json,sys; d=json.loads(open('$O/bench_batch1.json').read()); print('batch 1 ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'], d['roofline']['avg_launch_ms'])"
tail -2 $O/bench_batch1.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_batch0.json 2> $O/bench_batch0.err
python -c "import json,sys; d=json.loads(open('$O/bench_batch0.json').read()); print('batch 0 (4-way reduction passes) ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'], d['roofline']['avg_launch_ms'])"
for c in 18 19; do
ZKB_MSM_C_G2=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_g2c$c.json 2> $O/bench_g2c$c.err
python -c "import json,sys; d=json.loads(open('$O/bench_g2c$c.json').read()); print('G2 c=$c ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'], d['roofline']['avg_launch_ms'])"
done
