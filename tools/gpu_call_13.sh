#!/bin/bash
# multi-GPU bench line: weak-scaling headline + sharded proof / sharded MSM / sharded Marlin sub-records
N=$1
O=gpurun_out/r2m_${N}gpu
mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | head -8 > $O/gpus.txt
( time NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
tail -c 3000 $O/bench.json; tail -8 $O/bench.err
