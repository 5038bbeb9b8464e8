#!/usr/bin/env python
"""Instruction counts per kernel of libdraupnir_b200.so from `cuobjdump -sass` (the SASS tells of tcgen05 / TMA / fp64 tensor
instructions):  python profiles/sass_counts.py > profiles/rNN/sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "draupnir_asr_b200", "libdraupnir_b200.so")
COLS = ["UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "LDGSTS", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "MUFU", "PRMT", "LDG", "STG", "LDS", "STS",
        "ELECT", "UTCATOM"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
rows, name, cnt, n = [], None, None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, n, cnt))
        name, cnt, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        n += 1
        op = m.group(2)
        for c in COLS:
            if op.startswith(c):
                cnt[c] += 1
if name:
    rows.append((name, n, cnt))
def short(mangled):
    out = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"\(anonymous namespace\)::", "", out)
    return re.sub(r"\(.*", "", out).replace("void ", "")
print("# SASS evidence, libdraupnir_b200.so (sm_100a): `cuobjdump -sass`, instruction counts per kernel (profiles/sass_counts.py)")
print("# tcgen05.mma -> UTCIMMA (kind::i8), tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async.bulk (TMA, non-tensor-map) -> UBLKCP,")
print("# cp.async -> LDGSTS, mbarrier -> SYNCS, mma.sync.m8n8k4.f64 -> DMMA, elect.sync -> ELECT, byte permute -> PRMT\n")
print("%-34s %7s " % ("kernel", "instrs") + " ".join("%7s" % c for c in COLS))
for name, n, cnt in sorted(rows, key=lambda r: short(r[0])):
    print("%-34s %7d " % (short(name)[:34], n) + " ".join("%7d" % cnt[c] for c in COLS))
tot = collections.Counter()
for _, _, cnt in rows:
    tot.update(cnt)
print("\n%-34s %7d " % ("library total", sum(r[1] for r in rows)) + " ".join("%7d" % tot[c] for c in COLS))
