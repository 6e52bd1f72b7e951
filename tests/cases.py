"""Shared definitions of the BASELINE.json configurations for tests, golden
generation and bench.py."""
import numpy as np

# name: (tiles_x, tiles_y, grid rows, grid cols, fov, frame_len N, history H, window W)
CONFIGS = {
    "cfg1": dict(nx=1, ny=1, rows=100, cols=100, fov=180.0, N=256, H=256, W=1024),
    "cfg2": dict(nx=4, ny=1, rows=64, cols=64, fov=180.0, N=256, H=256, W=1024),
    "cfg3": dict(nx=4, ny=2, rows=32, cols=32, fov=180.0, N=256, H=256, W=1024),
    "cfg5": dict(nx=4, ny=2, rows=256, cols=256, fov=180.0, N=4096, H=256, W=4608),
}
# cfg4: 512 mics, 16 tracked targets on a ring theta = 30 deg, phi = k * 22.5 deg
CFG4 = dict(nx=4, ny=2, T=16, N=256, H=256, W=1024)


def cfg4_targets():
    k = np.arange(CFG4["T"])
    return np.full(CFG4["T"], np.deg2rad(30.0)), np.deg2rad(22.5) * k


def origins(nx, ny, pitch=0.16):
    o = [[(i - (nx - 1) / 2) * pitch, (j - (ny - 1) / 2) * pitch, 0.0] for j in range(ny) for i in range(nx)]
    return np.asarray(o, np.float32)
