#!/usr/bin/env python
"""Single-frame kernel time of one configuration for forced latency shapes (BFLK_LAT_WARPS x BFLK_LAT_SPLIT):
usage: python tools/latency_sweep.py cfg3 16x8 16x4 8x4 ...   (each shape runs in its own process: the knobs are read once)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, ctypes as C, os
ROOT = sys.argv[1]; name = sys.argv[2]
sys.path[:0] = [ROOT, os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, bflk, cases
from bflk import synth
c = cases.CONFIGS[name]
w = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), c["rows"], c["cols"], c["fov"])
w.set_channel_split(True)
win = synth.make_stream(synth.tile_geometry(cases.origins(c["nx"], c["ny"])), 1024)
pin = torch.from_numpy(win).pin_memory(); out = torch.empty(c["rows"] * c["cols"], dtype=torch.float32).pin_memory()
def call():
    assert w._L.bflk_power_map(w._h, C.c_void_p(pin.data_ptr()), C.c_void_p(out.data_ptr())) == 0
for _ in range(30): call()
w.enable_timing(True); w.kernel_time_ms(); t = []
for _ in range(200):
    t0 = time.perf_counter(); call(); t.append((time.perf_counter() - t0) * 1e6)
das_ms, das_n, pack_ms, pack_n = w.kernel_time_ms()
print(f"{name} warps x split {os.environ.get('BFLK_LAT_WARPS')}x{os.environ.get('BFLK_LAT_SPLIT')}: p50 {np.percentile(t, 50):.0f} us, kernel {das_ms / das_n * 1e3:.1f} us, pack {pack_ms / max(1, pack_n) * 1e3:.1f} us")
'''
name = sys.argv[1]
for shape in sys.argv[2:]:
    wv, sv = shape.split("x")
    env = dict(os.environ, BFLK_LAT_WARPS=wv, BFLK_LAT_SPLIT=sv)
    subprocess.run([sys.executable, "-c", CHILD, ROOT, name], env=env)
