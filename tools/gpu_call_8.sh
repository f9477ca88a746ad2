#!/bin/bash
# batched-affine accumulation over equal-length chains (msm_batch.cuh v2): parity, stand-alone timings, sweep of lmax / blocks per SM
O=gpurun_out/r2h
mkdir -p $O
( time ZKB_MSM_BATCH=1 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu ) > $O/pytest_msm.log 2>&1
tail -5 $O/pytest_msm.log
timeout 300 python tools/exp_pair.py --levels 0 --batch 0 1 --steps 5 > $O/exp_batch.jsonl 2> $O/exp_batch.err
cat $O/exp_batch.jsonl; tail -3 $O/exp_batch.err
for bps in 3 4 5; do for lmax in 8 12 16; do
  echo "bps $bps lmax $lmax"
  ZKB_BATCH_BPS=$bps ZKB_BATCH_LMAX=$lmax timeout 300 python tools/exp_pair.py --levels 0 --batch 1 --groups 1 --steps 3 2>&1 | tail -1
done; done
for bt in 1; do
  ZKB_MSM_BATCH=$bt timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_batch$bt.json 2> $O/bench_batch$bt.err
  python -c "import json,sys; d=json.loads(open('$O/bench_batch$bt.json').read()); print('batch $bt ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'], d['roofline']['avg_launch_ms'])"
  tail -2 $O/bench_batch$bt.err
done
