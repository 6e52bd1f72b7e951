#!/usr/bin/env python
"""Where a single-frame bflk_power_map call spends its time: pageable vs page-locked input, device time of the pack
pre-pass and of the delay-and-sum kernel (CUDA events inside the library), rest = copies, launches, synchronisation."""
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
    for kernel, split in ((0, True), (0, False), (2, False)):
        w = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), c["rows"], c["cols"], c["fov"])
        w.set_kernel(kernel)
        w.set_channel_split(split)
        win = synth.make_stream(synth.tile_geometry(cases.origins(c["nx"], c["ny"])), 1024)
        pin = torch.from_numpy(win).pin_memory()
        out = torch.empty(c["rows"] * c["cols"], dtype=torch.float32).pin_memory()
        res = {}
        for label, src in (("pageable", win.ctypes.data), ("pinned", pin.data_ptr())):
            def call():
                rc = w._L.bflk_power_map(w._h, C.c_void_p(src), C.c_void_p(out.data_ptr()))
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
            res[label] = (np.percentile(t, 50), np.percentile(t, 95), das_ms / das_n * 1e3, pack_ms / max(1, pack_n) * 1e3)
        k = w.kernel_info()
        print(f"{name} kernel {k[0]}{' + channel split' if split else ''}: pageable p50 {res['pageable'][0]:.0f} us (p95 {res['pageable'][1]:.0f}); page-locked p50 {res['pinned'][0]:.0f} us "
              f"(p95 {res['pinned'][1]:.0f}) of which delay-and-sum kernel {res['pinned'][2]:.0f} us, pack {res['pinned'][3]:.0f} us")
        w.close()
