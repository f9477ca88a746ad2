#!/bin/bash
# ncu full capture of the batched-affine level kernel (level 0 of a 2^20 G1 MSM) with source attribution
O=gpurun_out/r2d
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair_level -s 2 -c 1 -o $O/pair_level_ncu \
   python tools/exp_pair.py --levels 1 --scales 4 --groups 1 --steps 1 > $O/ncu_pair.log 2>&1
tail -3 $O/ncu_pair.log
ls -la $O
