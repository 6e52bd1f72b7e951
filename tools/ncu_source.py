#!/usr/bin/env python
"""Per-SASS-instruction stall samples of the hottest region of an .ncu-rep (first kernel).
usage: python tools/ncu_source.py report.ncu-rep [first_line last_line]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(body))
tot = sum(int(r[idx["# Samples"]]) for r in body)
print(f"total samples {tot}")
for n, r in enumerate(body[lo:hi], lo):
    s = int(r[idx["# Samples"]])
    top = sorted(((int(r[idx[k]]), k[6:]) for k in stalls), reverse=True)[:3]
    tops = " ".join(f"{k}:{v}" for v, k in top if v)
    print(f"{n:5d} {s:6d} {int(r[idx['Instructions Executed']]):10d}  {r[idx['Source']].strip():70s} {tops}")
