#!/bin/bash
# round-2 GPU call C: launch lists of the stand-alone G2 / G1 MSM (where do the 43 ms of the G2 XYZZ path go?), a proper
# capture of k_pair_level, initcheck, and the bench line with the Marlin sub-record.
O=gpurun_out/r2c
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_g2_msm.csv \
   python tools/exp_pair.py --levels 0 --groups 2 --steps 1 > $O/l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_g2_msm_lv2.csv \
   python tools/exp_pair.py --levels 2 --scales 4 --groups 2 --steps 1 > $O/l2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pair_level' -s 2 -c 2 -o $O/pair_level_ncu \
   python tools/exp_pair.py --levels 2 --scales 4 --groups 1 --steps 1 > $O/ncu_pair.log 2>&1
TOOLS=initcheck TMO=600 bash tools/sanitize.sh $O/sanitizer > $O/sanitize_summary.txt 2>&1
cat $O/sanitize_summary.txt
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
tail -c 1500 $O/bench.json; tail -5 $O/bench.err
