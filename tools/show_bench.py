#!/usr/bin/env python
"""Prints the interesting keys of a bench.py JSON line (file argument or stdin)."""
import json
import sys
l = json.loads((open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin).read().strip().splitlines()[-1])
for k in ("impl", "value", "ms_per_step", "n_gpus", "gpu_launches"):
    if k in l:
        print(k, l[k])
print("config", {k: l["config"].get(k) for k in ("parallelism", "kernel", "frames_per_step", "directions_per_gpu", "frames_per_gpu")})
print("clocks", l.get("clocks"))
print("e2e", l.get("e2e"))
if "roofline" in l and isinstance(l["roofline"], dict):
    print("roofline", {k: l["roofline"].get(k) for k in ("kernel", "frac", "achieved", "ffma_ubench_tflops", "frac_vs_ffma_ubench", "kernel_share_of_step", "pack_share_of_step", "traffic")})
for k in ("bit_identical", "sustained", "grid_shard", "comm", "latency_single_frame_us", "cpu_baseline"):
    if k in l:
        print(k, l[k])
for k, v in l.get("other_configs", {}).items():
    print("other", k, {a: b for a, b in v.items() if a in ("value", "roofline_frac", "roofline_frac_rank0", "kernel", "error", "latency_resident_window", "latency_with_upload", "latency_monopulse_26_particles", "device_us_per_call", "parallelism")})
