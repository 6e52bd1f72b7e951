"""Regenerates tests/golden/*.npz from the CPU oracle and the compiled reference kernel (oracle/_ref).

Run in the build container (needs /root/reference for oracle/_ref):  python tests/golden/make_golden.py
The fixtures travel to the GPU box; nothing at test time reads /root/reference.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")]

from oracle import oracle as O  # noqa: E402
from bflk import synth  # noqa: E402
import cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert O.ref() is not None, "oracle/_ref/libref.so missing (needs /root/reference)"
    # ---- A: steering tables of every configuration (hash + sampled rows) ----
    out = {}
    for name, c in cases.CONFIGS.items():
        xyz = O.create_tiled_antenna(cases.origins(c["nx"], c["ny"]))
        off, fr = O.mimo_lut(xyz, c["rows"], c["cols"], c["fov"], c["H"])
        D = off.shape[0]
        rows = np.arange(0, D, max(1, D // 16))
        out[f"{name}_off_sha"] = np.frombuffer(bytes.fromhex(sha(off)), np.uint8)
        out[f"{name}_frac_sha"] = np.frombuffer(bytes.fromhex(sha(fr)), np.uint8)
        out[f"{name}_rows"] = rows
        out[f"{name}_off_rows"] = off[rows]
        out[f"{name}_frac_rows"] = fr[rows]
        out[f"{name}_max_delay"] = np.int32((c["H"] - off).max())
        print(name, "D", D, "C", xyz.shape[0], "max delay", (c["H"] - off).max())
    # odd grid with the degenerate centre cell
    xyz1 = O.create_antenna()
    off, fr = O.mimo_lut(xyz1, 9, 9, 120.0)
    out["odd9_off"], out["odd9_frac"] = off, fr
    out["antenna_xyz"] = xyz1
    # B: cfg4 dynamic steering tables
    th, ph = cases.cfg4_targets()
    xyz8 = O.create_tiled_antenna(cases.origins(4, 2))
    out["cfg4_off"], out["cfg4_frac"] = O.steer_tables(xyz8, th, ph)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **out)

    # ---- C/D/E/F: one seeded 8x8 snapshot through every function of the path ----
    xyz = O.create_antenna()
    window = synth.make_stream(xyz, 1024)                 # [64][1024]
    off, fr = O.mimo_lut(xyz, 16, 16, 180.0)
    power = O.mimo_update(window, off, fr)
    das = O.mimo_das(window, off, fr)
    ref_power, ref_das = O.ref_mimo_update(window, off, fr, want_das=True)
    assert np.array_equal(das, ref_das), "oracle delay-and-sum differs from the compiled reference delay()"
    mask = np.array([i for i in range(64) if i not in (3, 17, 40)], np.int32)
    power_masked = O.mimo_update(window, off, fr, index=mask)
    sel = np.array([0, 37, 128, 255])
    th = np.deg2rad([20.0, 45.0, 0.0, 63.0])
    ph = np.deg2rad([30.0, 200.0, 0.0, 310.0])
    soff, sfr = O.steer_tables(xyz, th, ph)
    audio = np.stack([O.particle_das(window, soff[t], sfr[t]) for t in range(4)])
    beam = np.array([O.particle_beam(window, soff[t], sfr[t]) for t in range(4)], np.float64)
    heat, arg, mx = O.populate_heatmap(power)
    cal = window.copy()
    cal[5] *= 0.0
    cal[9] *= 3.0
    cal[33] *= 0.5
    cidx, ccorr, cmed, cmean = O.calibrate(cal)
    wire = synth.to_wire_i32(window[:, :64])
    exposure = O.ingest(wire)
    np.savez_compressed(
        os.path.join(HERE, "snapshot.npz"), window=window, power=power, ref_power=ref_power, das_sel=sel,
        das=das[sel], power_masked=power_masked, mask=mask, miso_theta=th, miso_phi=ph, miso_off=soff, miso_frac=sfr,
        miso_audio=audio, miso_beam=beam, heat=heat, heat_argmax=np.int32(arg), heat_max=np.float32(mx),
        cal_index=cidx, cal_corr=ccorr, cal_median=np.float32(cmed), cal_mean=np.float32(cmean),
        wire=wire, exposure=exposure)
    print("power peak", power.argmax(), "ref/oracle power max rel diff", np.max(np.abs(ref_power - power) / power))

    # ---- G: the FIR variant of delay() (delay.cpp:28-40) as the reference file compiles it without AVX2 ----
    assert O.ref_fir() is not None, "oracle/_ref/libref_fir.so missing (needs /root/reference)"
    coeffs = O.ref_filter_table()                          # src/dsp/filter.h:10-112 as the compiled object holds it
    fir_ref_power = O.ref_mimo_update_fir(window, off, fr)
    fir_power = O.mimo_update_fir(window, off, fr, coeffs)
    rng = np.random.default_rng(5)
    sig = (0.02 * rng.standard_normal((8, 300))).astype(np.float32)
    fracs = rng.random(8).astype(np.float32)
    outs = np.zeros((8, 256), np.float32)
    for k in range(8):
        O.ref_fir().ref_delay(outs[k], sig[k], fracs[k])
    np.savez_compressed(os.path.join(HERE, "fir.npz"), coeffs=coeffs, ref_power=fir_ref_power, power=fir_power,
                        kat_signal=sig, kat_fraction=fracs, kat_out=outs)
    print("FIR: oracle vs compiled reference power max rel diff", np.max(np.abs(fir_ref_power - fir_power) / fir_power))
    # ---- H: cv::resize(INTER_LINEAR) on 8-bit maps, from the real OpenCV (opencv-python; the reference links libopencv) ----
    import cv2
    rng = np.random.default_rng(9)
    res = {"cv2_version": np.array(cv2.__version__)}
    heat16 = heat.reshape(16, 16)
    big = cv2.resize(heat16, (1024, 1024), interpolation=cv2.INTER_LINEAR)       # AWProcessingUnit::draw: compact -> 1024 x 1024
    res["heat16"], res["heat16_to_1024_sha"] = heat16, np.frombuffer(bytes.fromhex(sha(big)), np.uint8)
    res["heat16_to_1024_rows"] = big[::97]
    for k, (ih, iw, oh, ow) in enumerate(((16, 16, 37, 53), (100, 100, 1024, 1024), (9, 11, 64, 48), (256, 256, 1024, 1024), (100, 100, 50, 60), (1, 1, 8, 8))):
        src = rng.integers(0, 256, (ih, iw), dtype=np.uint8)
        out = cv2.resize(src, (ow, oh), interpolation=cv2.INTER_LINEAR)
        res[f"case{k}_src"] = src
        res[f"case{k}_shape"] = np.array([oh, ow])
        res[f"case{k}_sha"] = np.frombuffer(bytes.fromhex(sha(out)), np.uint8)
        if out.size <= 4096:
            res[f"case{k}_out"] = out
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **res)
    for f in ("tables.npz", "snapshot.npz", "fir.npz", "resize.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
