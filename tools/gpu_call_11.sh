#!/bin/bash
O=gpurun_out/r2k
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "batch" ) > $O/pytest_msm.log 2>&1
tail -5 $O/pytest_msm.log
ZKB_MSM_BATCH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_accumulate_chains|k_chain_combine' -s 6 -c 3 -o $O/chains3_ncu \
   python tools/exp_pair.py --levels 0 --batch 1 --groups 1 --steps 1 > $O/ncu_chains.log 2>&1
tail -3 $O/ncu_chains.log
