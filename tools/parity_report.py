#!/usr/bin/env python
"""Measured parity of the power maps against the CPU oracle on the BASELINE.json configurations (bar: 1e-4 max
relative error, identical peak direction).  Kernel 2 = exact operation triple, kernel 4 = two-FMA form (automatic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import bflk  # noqa: E402
from bflk import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
import cases  # noqa: E402

for name in ("cfg1", "cfg2", "cfg3", "cfg5"):
    c = cases.CONFIGS[name]
    org = cases.origins(c["nx"], c["ny"])
    w = bflk.MIMOWorker(org, c["rows"], c["cols"], c["fov"], frame_len=c["N"], history=c["H"], window_len=c["W"])
    window = synth.make_stream(synth.tile_geometry(org), c["W"])
    off, fr = w.tables()
    D = off.shape[0]
    sel = np.arange(D) if name != "cfg5" else np.unique(np.r_[0:8, np.arange(17, D, 1021), D - 8:D])
    po = O.mimo_update(window, off[sel], fr[sel], n=c["N"])
    for k in (2, 4):
        w.set_kernel(k)
        p = w.update(window)
        err = np.abs(p[sel].astype(np.float64) - po) / po
        print(f"{name} kernel {k}: max rel err {err.max():.2e}, median {np.median(err):.1e} over {len(sel)} directions, "
              f"peak direction {'identical' if name == 'cfg5' or int(np.argmax(p)) == int(np.argmax(po)) else 'DIFFERS'}"
              + (f" (argmax {int(np.argmax(p))})" if name != "cfg5" else ""))
