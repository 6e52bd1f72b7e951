"""The C++ drop-in boundary: beamforming-lk_b200/host/cuda_workers.h (CudaMIMOWorker / CudaMISOWorker, the
reference's Worker plugin interface) compiled against stand-in reference headers and driven like
AWProcessingUnit drives its workers."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT

SRC = os.path.join(ROOT, "tests", "standin", "adapter_main.cpp")


def build(tmp_path):
    exe = str(tmp_path / "adapter_main")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-DBFLK_STANDIN_HEADERS", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(PKG, "host"), "-I" + os.path.join(ROOT, "tests", "standin"), "-o", exe, SRC,
           "-L" + PKG, "-lbflk", "-Wl,-rpath," + PKG, "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapter_compiles_against_worker_interface(tmp_path):
    exe = build(tmp_path)
    # no GPU here: the adapter must report the failure (no CPU fallback) and still shut down cleanly
    import torch
    if not torch.cuda.is_available():
        stream = np.zeros((64, 1024), np.float32)
        path = str(tmp_path / "s.f32")
        stream.tofile(path)
        r = subprocess.run([exe, path, "1024", "4", "4", "180", "0", "0"], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0
        assert "no CPU fallback" in r.stderr or "no CUDA device" in r.stderr
        assert "power 0 0 0" in r.stdout          # powerdB untouched
        assert "deleted_through_base 1" in r.stdout


@pytest.mark.gpu
def test_adapter_matches_golden_on_gpu(tmp_path, golden):
    g = golden["snapshot"]
    exe = build(tmp_path)
    path = str(tmp_path / "s.f32")
    np.ascontiguousarray(g["window"], np.float32).tofile(path)
    th, ph = float(g["miso_theta"][0]), float(g["miso_phi"][0])
    r = subprocess.run([exe, path, "1024", "16", "16", "180", repr(th), repr(ph)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = {l.split(" ", 1)[0]: l.split(" ", 1)[1] for l in r.stdout.strip().splitlines()}
    assert out["type"] == "2 3"                   # worker_t::MIMO, worker_t::MISO
    power = np.array(out["power"].split(), np.float64)
    assert np.max(np.abs(power - g["power"]) / g["power"]) <= 1e-4
    assert int(np.argmax(power)) == int(np.argmax(g["power"]))
    heat = np.array(out["heat"].split(), np.int64)
    assert np.max(np.abs(heat - g["heat"].astype(np.int64))) <= 1 and heat.max() == 255
    audio = np.array(out["audio"].split(), np.float32)
    assert np.array_equal(audio, g["miso_audio"][0])          # das(): bit-exact (%.9g round-trips float32)
    assert abs(float(out["beam"]) - g["miso_beam"][0]) <= 1e-4 * g["miso_beam"][0]
    # Worker::tracking as TargetHandler reads it (AWProcessingUnit::targets): the map's peaks, per the oracle's definition
    from oracle import oracle as O
    idx, pw, pr = O.map_targets(g["power"], 16, 16, 8, 0.5)
    th_grid, ph_grid = O.mimo_grid(16, 16, 180.0)
    t = np.array(out["targets"].split(), np.float64).reshape(-1, 4)
    assert t.shape[0] == len(idx) >= 1
    assert np.array_equal(t[:, 0], th_grid[idx]) and np.array_equal(t[:, 1], ph_grid[idx])
    assert np.allclose(t[:, 2], pw, rtol=1e-4) and np.allclose(t[:, 3], pr, rtol=2e-3)
    assert out["start_kept"] == "1"               # the same target on the next frame keeps its first-found time
    assert out["deleted_through_base"] == "1"     # delete through Worker* (non-virtual ~Worker) shut down cleanly
