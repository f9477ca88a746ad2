#!/bin/bash
# timeline of ONE proof sharded over N ranks (rank 0's view)
N=$1
O=gpurun_out/r2t_${N}gpu
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/timeline.py --sharded --out $O/timeline_sharded.txt > $O/tl.log 2>&1
tail -5 $O/tl.log
head -40 $O/timeline_sharded.txt
