#!/bin/bash
# round 2, call 28: rank 0's share of an 8-way sharded 2^20 proof on one GPU (no exchange): timeline
O=gpurun_out/r2ae
mkdir -p $O
cd /root/repo
timeout 600 python tools/timeline.py --log-constraints 20 --partial-of 8 --out $O/timeline_partial8.txt > $O/tl.log 2>&1
tail -3 $O/tl.log | cut -c1-200
head -36 $O/timeline_partial8.txt | cut -c1-150
