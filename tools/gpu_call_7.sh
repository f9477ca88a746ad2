#!/bin/bash
# new parity tests (Groth16 setup on the GPU, C-ABI fixture, decompression), G2 accumulation with the multiplication as
# a call (instruction-fetch stalls), kernel timelines of the proof (default and with low-priority accumulation streams)
O=gpurun_out/r2g
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_generator.py tests/test_bindings.py tests/test_gpu_msm.py -x -q -m gpu ) > $O/pytest_new.log 2>&1
tail -8 $O/pytest_new.log
for call in 0 2 3; do
  ZKB_ACC_CALL=$call timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_call$call.json 2> $O/bench_call$call.err
  python -c "import json,sys; d=json.loads(open('$O/bench_call$call.json').read()); print('acc_call $call ms/proof', d['ms_per_step'], 'verified', d['verified_in_exponent'], d['roofline']['avg_launch_ms'])"
  tail -2 $O/bench_call$call.err
done
timeout 300 python tools/timeline.py --out $O/timeline_default.txt > $O/tl0.log 2>&1
ZKB_BULK=1 timeout 300 python tools/timeline.py --out $O/timeline_bulk.txt > $O/tl1.log 2>&1
head -4 $O/timeline_default.txt $O/timeline_bulk.txt
ZKB_BULK=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline > $O/bench_bulk.json 2> $O/bench_bulk.err
python -c "import json,sys; d=json.loads(open('$O/bench_bulk.json').read()); print('bulk ms/proof', d['ms_per_step'])"
