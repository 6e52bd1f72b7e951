"""The C-ABI shared library loads and exports every symbol include/bflk.h declares (no GPU needed),
and fails loudly -- no CPU fallback -- when no B200 is present."""
import ctypes
import os
import re
import subprocess

import pytest

import bflk
from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bflk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bflk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(bflk.SYMBOLS)


def test_library_exports_every_symbol():
    assert os.path.exists(bflk.LIB_PATH), "libbflk.so not built: run beamforming-lk_b200/build.sh"
    L = bflk.load_library()
    for name in header_symbols():
        assert hasattr(L, name), name
    exported = subprocess.run(["nm", "-D", "--defined-only", bflk.LIB_PATH], capture_output=True, text=True).stdout
    for name in header_symbols():
        assert re.search(rf"\bT {name}\b", exported), name
    assert L.bflk_version() == 1


def test_library_is_built_for_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", bflk.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_default_config_matches_reference_constants():
    L = bflk.load_library()
    cfg = bflk.Config()
    L.bflk_default_config(ctypes.byref(cfg))
    # streams.hpp:28-32, antenna.h:16-20
    assert (cfg.n_channels, cfg.frame_len, cfg.history, cfg.window_len) == (64, 256, 256, 1024)
    assert cfg.sample_rate == 48828.0 and cfg.propagation_speed == 340.0


def test_create_rejects_bad_config():
    L = bflk.load_library()
    cfg = bflk.Config()
    L.bflk_default_config(ctypes.byref(cfg))
    cfg.window_len = 300            # cannot hold history + frame + 1
    h = ctypes.c_void_p()
    assert L.bflk_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"window_len" in L.bflk_last_error(None)
    assert not h.value


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bflk.BflkError) as e:
        bflk.Beamformer()
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
