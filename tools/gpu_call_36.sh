#!/bin/bash
# round 2, call 36: inversion-free Miller loop + sparse line products
O=gpurun_out/r2aj
mkdir -p $O
cd /root/repo
( time timeout 1500 python -m pytest tests/test_gpu_pairing.py tests/test_gpu_marlin_proof.py tests/test_gpu_plonk.py -x -q -m gpu ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
timeout 900 python tools/bench_pairing.py --batches 1 8192 32768 > $O/pairing_bench.jsonl 2> $O/pairing_bench.err
cat $O/pairing_bench.jsonl; tail -3 $O/pairing_bench.err
