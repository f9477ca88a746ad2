#!/bin/bash
O=gpurun_out/r2i
mkdir -p $O
ZKB_MSM_BATCH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_accumulate_chains|k_chain_combine' -s 4 -c 2 -o $O/chains_ncu \
   python tools/exp_pair.py --levels 0 --batch 1 --groups 1 --steps 1 > $O/ncu_chains.log 2>&1
tail -3 $O/ncu_chains.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_batch.csv \
   python tools/exp_pair.py --levels 0 --batch 1 --groups 1 --steps 1 > $O/l.log 2>&1
ls -la $O
