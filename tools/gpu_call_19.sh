#!/bin/bash
# window width vs MSM size for mid-size MSMs (the shards of a sharded proof, Marlin's commitments): BLS12-381 G1 2^L, G2 2^(L-1)
O=gpurun_out/r2u
mkdir -p $O
for L in 17 18 19; do for c in 15 16 17 18 19 20; do
  echo "log_n $L c $c"
  ZKB_MSM_C=$c timeout 300 python tools/exp_pair.py --log-n $L --levels 0 --batch 0 --steps 5 2>&1 | grep -o '"group": [12], "log_n": [0-9]*\|"ms_median": [0-9.]*' | paste - - | sed 's/"//g'
done; done
