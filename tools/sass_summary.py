#!/usr/bin/env python
"""Opcode summary of the shipped kernels (cuobjdump -sass on libbflk.so): per kernel the instruction count and the
mnemonics that prove what the hot path is made of -- FFMA2 / FADD2 (packed FP32, sm_100-only), UBLKCP (TMA bulk copy),
SYNCS (mbarrier), LDS.128, and the absence of HMMA / UTCMMA (no tensor-core instruction: the path is a gather-FMA).
usage: python tools/sass_summary.py [kernel-name-regex] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "beamforming-lk_b200", "libbflk.so")
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name = None
per = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and name:
        op = m.group(1)
        if op in ("LDS", "LDG", "STG", "STS") and m.group(2):
            w = re.search(r"\.(64|128|256)", m.group(2))
            op += "." + (w.group(1) if w else "32")
        per[name][op] += 1
KEY = ["FFMA2", "FADD2", "FFMA", "FADD", "LDS.128", "LDS.64", "LDS.32", "LDG.32", "UBLKCP", "SYNCS", "BRA", "LOP3", "IADD3", "ATOMS", "UCGABAR_ARV", "MAPA", "HMMA", "UTCMMA", "UTMALDG"]
tot = collections.Counter()
print(f"SASS opcode summary of {os.path.relpath(LIB, ROOT)} (sm_100a); columns: total instructions, then selected mnemonics")
print(f"{'kernel':78s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in KEY))
for n, c in per.items():
    if pat and not pat.search(n):
        continue
    short = re.sub(r"bflk::|\(anonymous namespace\)::", "", n)
    short = re.sub(r"\(.*\)$", "", short)
    print(f"{short[:78]:78s} {sum(c.values()):7d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEY))
    tot.update(c)
print(f"{'ALL KERNELS IN THE LIBRARY':78s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(k, 0):7d}" for k in KEY))
print("tensor-core mnemonics (HMMA / UTCMMA / tcgen05): " + str(sum(v for k, v in tot.items() if k.startswith(("HMMA", "UTCMMA", "UTCHMMA", "IMMA")))))
