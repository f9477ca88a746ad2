#!/bin/bash
# batched-affine accumulation (msm_batch.cuh): parity, stand-alone MSM timings against the XYZZ loop, whole proof
O=gpurun_out/r2e
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_field.py -x -q -m gpu ) > $O/pytest_msm.log 2>&1
tail -5 $O/pytest_msm.log
timeout 600 python tools/exp_pair.py --levels 0 --batch 0 1 --steps 5 > $O/exp_batch.jsonl 2> $O/exp_batch.err
cat $O/exp_batch.jsonl; tail -3 $O/exp_batch.err
for bt in 0 1; do
  ZKB_MSM_BATCH=$bt timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_batch$bt.json 2> $O/bench_batch$bt.err
  python -c "import json,sys; d=json.loads(open('$O/bench_batch$bt.json').read()); print('batch $bt ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'], d['roofline']['avg_launch_ms'])"
  tail -2 $O/bench_batch$bt.err
done
