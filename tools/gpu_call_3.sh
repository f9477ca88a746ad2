#!/bin/bash
# round-2 GPU call C: full GPU test suite (with the decompression / C-ABI fixture tests), launch lists of the stand-alone
# G2 / G1 MSM, the bench line with the Marlin sub-record, and the ncu launch list of the bench command.
O=gpurun_out/r2c
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_g2_msm.csv \
   python tools/exp_pair.py --levels 0 --groups 2 --steps 1 > $O/l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_g2_msm_lv2.csv \
   python tools/exp_pair.py --levels 2 --scales 4 --groups 2 --steps 1 > $O/l2.log 2>&1
( time timeout 1200 python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
tail -c 2500 $O/bench.json; tail -5 $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python tools/prove_once.py --log-constraints 20 --proofs 2 > $O/prove_once.log 2>&1
N=$(grep PROOF_LAUNCHES $O/prove_once.log | tail -1 | awk '{print $2}')
python tools/launch_summary.py $O/launches.csv $N > $O/launches_summary.txt 2>&1
head -40 $O/launches_summary.txt
