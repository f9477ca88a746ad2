#!/bin/bash
# round 2, call 30: batched Groth16 verification by random linear combination
O=gpurun_out/r2ag
mkdir -p $O
cd /root/repo
( time timeout 1500 python -m pytest tests/test_gpu_pairing.py tests/test_bindings.py -x -q -m gpu ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
timeout 900 python tools/bench_pairing.py --batches 8192 > $O/pairing_bench.jsonl 2> $O/pairing_bench.err
cat $O/pairing_bench.jsonl; tail -3 $O/pairing_bench.err
