#!/usr/bin/env python
"""Single-frame bflk_power_map latency per kernel (2 exact tiled, 3 lane-broadcast, 4 two-FMA tiled), page-locked input."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bflk  # noqa: E402
from bflk import synth  # noqa: E402
import cases  # noqa: E402

for name in ("cfg1", "cfg2", "cfg3"):
    c = cases.CONFIGS[name]
    ref = None
    for kernel in (2, 3, 4):
        w = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), c["rows"], c["cols"], c["fov"])
        w.set_kernel(kernel)
        win = synth.make_stream(synth.tile_geometry(cases.origins(c["nx"], c["ny"])), 1024)
        pin = torch.from_numpy(win).pin_memory()
        out = torch.empty(c["rows"] * c["cols"], dtype=torch.float32).pin_memory()

        def call():
            rc = w._L.bflk_power_map(w._h, C.c_void_p(pin.data_ptr()), C.c_void_p(out.data_ptr()))
            assert rc == 0
        for _ in range(30):
            call()
        w.enable_timing(True)
        w.kernel_time_ms()
        t = []
        for _ in range(200):
            t0 = time.perf_counter()
            call()
            t.append((time.perf_counter() - t0) * 1e6)
        das_ms, das_n, pack_ms, pack_n = w.kernel_time_ms()
        w.enable_timing(False)
        o = out.numpy().copy()
        if ref is None:
            ref = o
        err = float(np.max(np.abs(o - ref) / np.maximum(np.abs(ref), 1e-30)))
        print(f"{name} kernel {kernel} (ran {w.kernel_info()[0]}): p50 {np.percentile(t, 50):.0f} us p95 {np.percentile(t, 95):.0f} us; "
              f"delay-and-sum kernel {das_ms / max(1, das_n) * 1e3:.0f} us, pack {pack_ms / max(1, pack_n) * 1e3:.0f} us; max rel diff vs kernel 2 {err:.2e}", flush=True)
        w.close()
