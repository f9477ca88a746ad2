#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals over the
last N launches (N = kernels per proof, printed by tools/prove_once.py as PROOF_LAUNCHES)."""
import collections
import csv
import re
import sys

path, last_n = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
        if r.get("Metric Name") == "gpu__time_duration.sum"]
last = rows[-last_n:]
tot = sum(v for _, v in last)
agg = collections.OrderedDict()
for k, v in last:
    k = re.sub(r"\(.*", "", k).replace("void ", "")
    a = agg.setdefault(k, [0.0, 0])
    a[0] += v
    a[1] += 1
print("launches: %d   serialized total: %.3f ms" % (len(last), tot / 1e6))
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-52s n=%3d %9.3f ms %5.1f%%" % (k[:52], c, v / 1e6, 100 * v / tot))
