#!/usr/bin/env python3
"""SASS opcode histogram of the hot kernels (no GPU needed): cuobjdump -sass on the per-TU objects, grouped per kernel.
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
Evidence for DESIGN.md's claim that the kernels are straight-line IMAD.WIDE code without local-memory spills."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "ckb_zkp_b200", "csrc", "build")
WANT = [("group_bls_g1.o", r"k_accumulate<.*BlsFq.*1, false"), ("group_bls_g2.o", r"k_accumulate<.*Fp2.*1, false"),
        ("group_bls_g1.o", r"k_pair_level"), ("group_bls_g1.o", r"k_seg_sum"), ("group_bls_g1.o", r"k_digits<true"),
        ("ntt.o", r"k_ntt_pass<.*Bls"), ("groth16_bls.o", r"k_spmv"), ("groth16_bls.o", r"k_g16_fold_finish")]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


cache = {}
for obj, pat in WANT:
    if obj not in cache:
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        funcs, cur = {}, None
        for line in txt.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                funcs[cur] = []
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                funcs[cur].append(m.group(2))
        cache[obj] = (funcs, demangle(list(funcs)))
    funcs, names = cache[obj]
    for mangled, ops in funcs.items():
        nice = names.get(mangled, mangled)
        if not re.search(pat, nice) or not ops:
            continue
        hist = collections.Counter(ops)
        base = collections.Counter(o.split(".")[0] for o in ops)
        print("=" * 110)
        print("%s  (%s)" % (nice[:200], obj))
        print("instructions: %d   IMAD.WIDE*: %d (%.1f %%)   local memory LDL/STL: %d / %d" % (
            len(ops), sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE")),
            100.0 * sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE")) / len(ops), base.get("LDL", 0), base.get("STL", 0)))
        print("  " + "  ".join("%s:%d" % kv for kv in base.most_common(24)))
        print("  async / bulk copies: " + (", ".join("%s:%d" % (k, v) for k, v in hist.items()
                                                    if k.startswith(("LDGSTS", "UTMALDG", "UBLKCP", "UTMASTG"))) or "none"))
        break
