#!/bin/bash
# round-2 final tree: full GPU suite, sanitizers over the extended target, bench line, launch list, ncu of the top kernels
O=gpurun_out/r2r
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
TOOLS="memcheck synccheck" TMO=900 bash tools/sanitize.sh $O/sanitizer > $O/sanitize_summary.txt 2>&1
TOOLS="racecheck" SMALL=1 TMO=1200 bash tools/sanitize.sh $O/sanitizer >> $O/sanitize_summary.txt 2>&1
cat $O/sanitize_summary.txt
( time timeout 1200 python bench.py --steps 20 --warmup 3 ) > $O/bench.json 2> $O/bench.err
python -c "import json; d=json.loads(open('$O/bench.json').read()); print('ms/proof', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'marlin', d['marlin']['ms_per_proof'], 'msm', d['msm']['ms_per_msm'], [ (x['field'],x['log_n'],round(x['fft']['ms'],3)) for x in d['ntt']['sizes']])"
tail -3 $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python tools/prove_once.py --log-constraints 20 --proofs 2 > $O/prove_once.log 2>&1
N=$(grep PROOF_LAUNCHES $O/prove_once.log | tail -1 | awk '{print $2}')
python tools/launch_summary.py $O/launches.csv $N > $O/launches_summary.txt 2>&1
head -14 $O/launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_accumulate|k_ntt_pass|k_seg_sum' -s 40 -c 14 -o $O/top_ncu \
   python tools/prove_once.py --log-constraints 20 --proofs 2 > $O/ncu_top.log 2>&1
tail -2 $O/ncu_top.log
timeout 600 python tools/bench_plonk.py --log-n 16 > $O/plonk16.json 2> $O/plonk16.err; cat $O/plonk16.json
timeout 600 python tools/bench_plonk.py --log-n 18 > $O/plonk18.json 2> $O/plonk18.err; cat $O/plonk18.json
