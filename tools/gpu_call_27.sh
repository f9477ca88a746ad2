#!/bin/bash
# round 2, call 27: what a rank of the 8-way sharded proof runs (MSM sizes of a 2^17 proof): timeline, G2 window widths
O=gpurun_out/r2ad
mkdir -p $O
cd /root/repo
timeout 300 python tools/timeline.py --log-constraints 17 --out $O/timeline_2e17.txt > $O/tl.log 2>&1
head -45 $O/timeline_2e17.txt | cut -c1-150
for c in 14 15 16 17 18; do
  echo "G2 c=$c"; ZKB_MSM_C_G2=$c timeout 200 python tools/exp_pair.py --log-n 18 --levels 0 --batch 0 --groups 2 --scales 0 2>&1 | grep ms_median
done
