#!/bin/bash
# round 2, call 38: window width vs size for BN254 G1 MSMs with resident tables (the Marlin commitments)
O=gpurun_out/r2al
mkdir -p $O
cd /root/repo
for L in 18 19 20; do
  for c in 16 17 18 20; do
    echo -n "bn254 log_n $L c $c  "
    ZKB_MSM_C=$c timeout 200 python tools/exp_pair.py --curve 0 --log-n $L --levels 0 --batch 0 --groups 1 --scales 0 2>&1 | grep -o '"ms_median": [0-9.]*'
  done
done | tee $O/bn254_window_sweep.txt
