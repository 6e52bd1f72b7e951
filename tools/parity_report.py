#!/usr/bin/env python
"""Measured parity of the power maps against the CPU reference path on the BASELINE.json configurations (bar: 1e-4 max
relative error, identical peak direction).  Kernel 2 = exact operation triple, kernel 4 = two-FMA form (automatic).

Three inputs per configuration: the SURVEY 8d signal (three tones + white noise, sigma = 1e-3), the same tones WITHOUT
noise (sigma = 0) and a single noise-free 3 kHz tone, where the side-lobe nulls are many orders of magnitude below the peak and any re-rounding of the
channel sum shows up in the relative error of those directions.  The CPU side is the compiled reference delay() loop
(oracle/_ref, all host threads) when it is available, else the C restatement (bit-identical delayed sums, see
tests/test_oracle.py); every direction of every grid is compared (cfg5: all 65 536).

usage: python tools/parity_report.py [cfg1 cfg2 ...] > profiles/rN_parity_report.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import bflk  # noqa: E402
from bflk import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
import cases  # noqa: E402


def cpu_power(window, off, fr, n):
    if O.ref() is not None:
        return O.ref_mimo_update(window, off, fr, n=n, n_threads=os.cpu_count() or 1), "compiled reference delay() loop"
    return O.mimo_update(window, off, fr, n=n), "oracle.c restatement"


def main():
    names = [a for a in sys.argv[1:] if a in cases.CONFIGS] or ["cfg1", "cfg2", "cfg3", "cfg5"]
    for name in names:
        c = cases.CONFIGS[name]
        org = cases.origins(c["nx"], c["ny"])
        w = bflk.MIMOWorker(org, c["rows"], c["cols"], c["fov"], frame_len=c["N"], history=c["H"], window_len=c["W"])
        off, fr = w.tables()
        one_tone = ((np.deg2rad(20.0), np.deg2rad(30.0), 3000.0, 1e-2),)
        for sigma, sources, label in ((1e-3, synth.DEFAULT_SOURCES, "3 tones"), (0.0, synth.DEFAULT_SOURCES, "3 tones"), (0.0, one_tone, "1 tone")):
            window = synth.make_stream(synth.tile_geometry(org), c["W"], sources=sources, sigma=sigma)
            po, who = cpu_power(window, off, fr, c["N"])
            po = po.astype(np.float64)
            for k in (2, 4):
                w.set_kernel(k)
                p = w.update(window).astype(np.float64)
                err = np.abs(p - po) / po
                worst = int(np.argmax(err))
                rel_to_peak = np.abs(p - po).max() / po.max()
                print(f"{name} {label} sigma {sigma:g} kernel {k}: max rel err {err.max():.2e} (direction {worst}, power {po[worst] / po.max():.1e} of the peak), "
                      f"median {np.median(err):.1e}, weakest direction {po.min() / po.max():.1e} of the peak, max |diff| / peak {rel_to_peak:.1e}, {len(po)} directions, "
                      f"peak direction {'identical' if int(np.argmax(p)) == int(np.argmax(po)) else 'DIFFERS'} ({int(np.argmax(po))}); CPU: {who}")
                sys.stdout.flush()


if __name__ == "__main__":
    main()
