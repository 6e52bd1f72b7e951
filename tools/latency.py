#!/usr/bin/env python
"""Single-call latencies of the host-buffer C-ABI entry points (what a live Worker::update() pays per frame):
bflk_power_map (one frame, cfg1 / cfg3) and bflk_miso (cfg4: 16 targets x 512 microphones, audio + beam power)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import bflk  # noqa: E402
from bflk import synth  # noqa: E402
import cases  # noqa: E402


def timeit(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    t = np.array(t) * 1e6
    return f"median {np.median(t):.0f} us, p95 {np.percentile(t, 95):.0f} us"


for name in ("cfg1", "cfg3"):
    c = cases.CONFIGS[name]
    w = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), c["rows"], c["cols"], c["fov"])
    win = synth.make_stream(synth.tile_geometry(cases.origins(c["nx"], c["ny"])), 1024)
    print(f"{name}: bflk_power_map, one 256-sample frame, {win.shape[0]} mics x {c['rows'] * c['cols']} directions: " + timeit(lambda: w.update(win)),
          "(real-time budget 5243 us per frame)")
c = cases.CFG4
m = bflk.MISOWorker(cases.origins(c["nx"], c["ny"]))
th, ph = cases.cfg4_targets()
m.steer(th, ph)
win = synth.make_stream(synth.tile_geometry(cases.origins(c["nx"], c["ny"])), 1024)
print("cfg4: bflk_miso, 16 targets x 512 mics, audio[16][256] + beam power: " + timeit(lambda: m.update(win)))
P = 26                                                           # 16 seekers + 10 trackers (gradient_ascend.h)
rng = np.random.default_rng(5)
pth, pph = rng.random(P) * np.deg2rad(70.0), rng.random(P) * 2 * np.pi
print("f2: bflk_monopulse, 26 particles x 4 beams x 512 mics (quadrant directions, tables, beams, gradient): "
      + timeit(lambda: m.monopulse(pth, pph, win, np.deg2rad(4.0), np.deg2rad(80.0), 3e-4)))
m.set_window(win)
print("cfg4: bflk_miso on the RESIDENT window (bflk_set_window once per frame), 16 targets: " + timeit(lambda: m.miso(th, ph)))
print("f2: bflk_monopulse on the resident window, 26 particles x 4 beams: "
      + timeit(lambda: m.monopulse(pth, pph, None, np.deg2rad(4.0), np.deg2rad(80.0), 3e-4)))
