#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel) into the handful of metrics DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py report.ncu-rep [--stalls]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active % (FP32 FMA-pipe utilisation)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe instructions %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe instructions %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
    ("smsp__warps_active.avg.per_cycle_active", "active warps / scheduler"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "achieved HBM read rate"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"kernel: {r[idx['Kernel Name']]}")
    for k, label in KEYS:
        if k in idx:
            print(f"  {label:58s} {r[idx[k]]} {units[idx[k]]}")
    if "--stalls" in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        his = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
        hi, end = his[0], (his[1] - 1 if len(his) > 1 else len(rows))
        h = rows[hi]
        ix = {n: i for i, n in enumerate(h)}
        cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot, ex, st = collections.Counter(), collections.Counter(), collections.Counter()

        def toi(x):
            try:
                return int(x)
            except ValueError:
                return 0
        for x in rows[hi + 1:end]:
            if len(x) < len(h):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", x[ix["Source"]])
            if not m:
                continue
            tot[m.group(2)] += toi(x[ix["# Samples"]])
            ex[m.group(2)] += toi(x[ix["Instructions Executed"]])
            for c in cols:
                st[c] += toi(x[ix[c]])
        S, E = sum(tot.values()), sum(ex.values())
        print("  warp-stall samples: " + ", ".join(f"{k[6:]} {100 * v / S:.1f}%" for k, v in st.most_common(9)))
        print("  instruction mix:    " + ", ".join(f"{k} {100 * v / E:.1f}%" for k, v in ex.most_common(10)))


if __name__ == "__main__":
    main()
