#!/bin/bash
# round-2 GPU call A: parity suite, bench (both arms), tuning sweeps, sanitizer.  Everything lands in gpurun_out/r2a/.
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
tail -c 600 $O/bench.json
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 0 ) > $O/bench_ref.json 2> $O/bench_ref.err
tail -c 400 $O/bench_ref.json
timeout 600 python tools/exp_pair.py > $O/exp_pair.jsonl 2>&1
for tile in 10 11; do for run in 1 2; do
  ZKB_NTT_TILE=$tile ZKB_NTT_RUN=$run timeout 300 python tools/bench_ntt.py --min-log 18 --max-log 24 --steps 5 > $O/ntt_t${tile}_r${run}.jsonl 2>&1
done; done
TOOLS="memcheck racecheck" TMO=420 bash tools/sanitize.sh $O/sanitizer > $O/sanitize_summary.txt 2>&1
cat $O/sanitize_summary.txt
