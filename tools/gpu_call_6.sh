#!/bin/bash
# ncu full capture of the batched-affine accumulation kernel (2^20 G1 MSM) with source attribution
O=gpurun_out/r2f
mkdir -p $O
ZKB_MSM_BATCH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate_batch -s 2 -c 1 -o $O/batch_ncu \
   python tools/exp_pair.py --levels 0 --batch 1 --groups 1 --steps 1 > $O/ncu_batch.log 2>&1
tail -3 $O/ncu_batch.log
ls -la $O
