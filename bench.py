#!/usr/bin/env python
"""bench.py -- full-grid delay-and-sum power maps per second on B200 (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--frames B] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of B synthetic frames (C channels x (B+3)*N samples,
larger than L2) for every direction of the grid.  N = 1 runs the configuration the target is quoted on
(cfg3: 512 microphones x 32x32 = 1024 directions x 256-sample frames).  N > 1 (torchrun, one rank per GPU)
arranges the ranks as G_d direction groups x G_f frame groups (default G_d = 2; --dir-groups N = the grid sharded
N ways): rank r computes direction slice r % G_d of the steering grid for frame slice r // G_d of the batch --
total work fixed, "strong" scaling -- and one NCCL all-gather inside the timed region assembles the [B][D] maps
on every rank.

One JSON line on rank 0.  `value`: maps/s with inputs resident in HBM.  `e2e`: the same through the
host-buffer C-ABI call (H2D of the batch + D2H of the maps inside the timed region).  `roofline`: the
dominant kernel (das_tile) against the FP32-FMA roofline the path is bound by, `roofline_hbm` against the
measured copy bandwidth.  `cpu_baseline`: the compiled reference delay() loop (oracle/_ref) on the host.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import cases  # noqa: E402

METRIC = "power_maps_per_sec"
UNIT = "maps/s"
FP32_LANES_PER_SM = 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=list(cases.CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step (default: sized so the input exceeds L2)")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 register-tiled (bit-identical sums), 3 lane-broadcast, 4 register-tiled two-FMA form")
    ap.add_argument("--dir-groups", type=int, default=0,
                    help="N > 1: direction groups G_d (ranks = G_d x frame groups); 0 = 2 when N is even; N = pure grid sharding")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def default_frames(c):
    # input stream = C x (B + 3) x N floats, at least 256 MB (L2 is 126 MB).  The kernel works on pairs of 256-sample
    # blocks, one CTA per (16 direction tiles, block pair): B is rounded up so that the number of block pairs is a
    # multiple of the 148 SMs -- the CTAs then fill whole waves at 1, 2, 4 and 8 GPUs (cfg3: B = 592, 310 MB)
    C = 64 * c["nx"] * c["ny"]
    per_frame = C * c["N"] * 4
    B = max(2, int(np.ceil(256e6 / per_frame)))
    if c["N"] == 256:
        pairs = -(-B // 2)
        B = 2 * (-(-pairs // 148) * 148)
    return B


def workload(c, name, B):
    C = 64 * c["nx"] * c["ny"]
    D = c["rows"] * c["cols"]
    T = (B - 1) * c["N"] + c["W"]
    T += T & 1
    return dict(workload=f"{name}: {C} mics x {c['rows']}x{c['cols']}={D} directions x {c['N']}-sample frames",
                channels=C, directions=D, frame_len=c["N"], frames_per_step=B, samples_per_channel=T,
                input_mb=round(C * T * 4 / 1e6, 1), l2="inputs larger than L2 (126 MB); no flush needed",
                frames_rule="256 MB of input, block pairs rounded up to a multiple of 148 SMs (whole CTA waves at 1-8 GPUs)")


def flops_per_map(C, D, N):
    return D * N * (4 * C + 6)          # 4 FLOP per (dir, ch, sample) + 6 per (dir, sample)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop = index, [], set(), threading.Event()
        self.max_mhz = None
        self.thread = None
        self.t0 = self.t1 = None     # the timed region; only samples inside it are reported

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            while not self.stop.is_set():
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), mhz, r, names))
                time.sleep(0.005)
        except Exception as e:  # NVML unavailable: report that rather than fail the bench
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=2)

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def summary(self):
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or 1e30)]
        for _, _, r, names in inside:
            for bit, nm in names.items():
                if r & bit:
                    self.reasons.add(nm)
        return {"sm_mhz": statistics.median([s[1] for s in inside]) if inside else None, "sm_max_mhz": self.max_mhz,
                "samples": len(inside), "reasons": sorted(self.reasons)}


def make_input(c, T):
    from bflk import synth
    xyz = synth.tile_geometry(cases.origins(c["nx"], c["ny"]))
    # tones are cheap, white noise for 512 x 132k samples is not: tile a 16-frame noise block
    base = synth.make_stream(xyz, T, sigma=0.0)
    rng = np.random.Generator(np.random.Philox(synth.SEED))
    blk = (1e-3 * rng.standard_normal((xyz.shape[0], 16 * c["N"]))).astype(np.float32)
    reps = int(np.ceil(T / blk.shape[1]))
    base += np.tile(blk, (1, reps))[:, :T]
    return np.ascontiguousarray(base, np.float32)


def cpu_reference_rate(c, stream, n_threads, budget_s, frames=None):
    """maps/s of the compiled reference delay() loop (oracle/_ref) over all directions, n_threads threads."""
    from oracle import oracle as O
    if O.ref() is None:
        raise RuntimeError("oracle/_ref/libref.so is missing")
    xyz = O.create_tiled_antenna(cases.origins(c["nx"], c["ny"]))
    off, fr = O.mimo_lut(xyz, c["rows"], c["cols"], c["fov"], c["H"])
    D = off.shape[0]
    sub = 1
    est_units = D * xyz.shape[0] * c["N"]
    if est_units / (1.5e9 * n_threads) > budget_s / 3:      # cfg5: time a direction subset and scale
        sub = int(np.ceil(est_units / (1.5e9 * n_threads) / (budget_s / 3)))
        off, fr = np.ascontiguousarray(off[::sub]), np.ascontiguousarray(fr[::sub])
    window = np.ascontiguousarray(stream[:, :c["W"]])
    O.ref_mimo_update(window, off, fr, n=c["N"], n_threads=n_threads)      # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        b = done % max(1, (stream.shape[1] - c["W"]) // c["N"] + 1)
        window = np.ascontiguousarray(stream[:, b * c["N"]: b * c["N"] + c["W"]])
        O.ref_mimo_update(window, off, fr, n=c["N"], n_threads=n_threads)
        done += 1
        el = time.perf_counter() - t0
        if (frames and done >= frames) or (not frames and el > budget_s):
            break
    rate = done / el / sub
    sample = f"{done} frames x {off.shape[0]} of {D} directions in {el:.2f} s" + (f" (1/{sub} direction subset, scaled)" if sub > 1 else "")
    return rate, sample


def run_reference(args, c, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.frames or default_frames(c)
    n_threads = os.cpu_count() or 1
    T = 15 * c["N"] + c["W"]
    stream = make_input(c, T)
    frames_per_step = 32      # ~0.13 s of all-core work per step at cfg3: a bounded sample of the B-frame step
    for _ in range(args.warmup):
        cpu_reference_rate(c, stream, n_threads, 1e9, frames=1)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, sample = cpu_reference_rate(c, stream, n_threads, 1e9, frames=frames_per_step)
        rates.append(r)
    el = time.perf_counter() - t0
    value = statistics.median(rates)
    cfg = workload(c, name, B)
    cfg["parallelism"] = f"{n_threads} host threads over directions"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "reference",
                             "sample": f"each step = {frames_per_step} frames of the workload; last: {sample}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "dir_samples_per_sec": value * cfg["directions"] * c["N"]}
    print(json.dumps(line))


def run_ours(args, c, name):
    import torch
    import torch.distributed as dist
    import bflk
    from bflk import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line and nothing else
        dist.init_process_group("nccl", device_id=dev)

    B = args.frames or default_frames(c)
    C, D, N = 64 * c["nx"] * c["ny"], c["rows"] * c["cols"], c["N"]
    cfg = workload(c, name, B)
    T = cfg["samples_per_channel"]
    # ranks = G_d direction groups x G_f frame groups (bflk/shard.py): rank r -> direction slice r % G_d of frame slice r // G_d
    gd, gf = shard.grid_2d(world, args.dir_groups) if world > 1 else (1, 1)
    dgrp, fgrp = rank % gd, rank // gd
    first, count = shard.direction_shard(D, gd, dgrp)
    per = shard.padded_count(D, gd)
    f0, nf = shard.frame_shard(B, gf, fgrp)
    nf_max = shard.frame_shard(B, gf, 0)[1]
    if nf <= 0:
        raise SystemExit(f"{B} frames do not fill {gf} frame groups")
    T_loc = (nf - 1) * N + c["W"]                          # this rank's slice of the stream: frames f0 .. f0 + nf

    w = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), c["rows"], c["cols"], c["fov"], device=local,
                        frame_len=N, history=c["H"], window_len=c["W"])
    w.set_kernel(args.kernel)
    w.set_direction_range(first, count)

    host_all = make_input(c, T)
    host_in = torch.from_numpy(np.ascontiguousarray(host_all[:, f0 * N: f0 * N + T_loc])).pin_memory()
    del host_all
    stream_dev = host_in.to(dev, non_blocking=True)
    local_pow = torch.zeros((nf_max, per), dtype=torch.float32, device=dev)
    tight = count == per and nf == nf_max                  # ragged shard: the kernel writes a [nf][count] buffer
    local_tight = local_pow if tight else torch.empty((nf, count), dtype=torch.float32, device=dev)
    gathered = torch.empty((world, nf_max, per), dtype=torch.float32, device=dev) if world > 1 else None
    host_out = torch.empty((B, D), dtype=torch.float32).pin_memory()
    in_group = None
    if world > 1 and gd > 1 and gf > 1:                    # ranks that share a frame slice replicate its input among themselves
        groups = [dist.new_group(list(range(g * gd, (g + 1) * gd))) for g in range(gf)]
        in_group = groups[fgrp]
    torch.cuda.synchronize()
    # a real (non-null) stream: kernels, NCCL and the timing events all go on it
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    cs = work_stream.cuda_stream
    assert cs != 0

    def step():
        w.power_map_batch_dev(stream_dev.data_ptr(), T_loc, nf, local_tight.data_ptr(), cs)
        if world > 1:
            if not tight:
                local_pow[:nf, :count].copy_(local_tight)
            dist.all_gather_into_tensor(gathered.view(world * nf_max, per), local_pow)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput, dominant-kernel time and clocks over the timed region ----
    with ClockSampler(local) as clk:       # NVML start-up overlaps the warm-up; samples are filtered to the timed region
        for _ in range(args.warmup):
            step()
        barrier()
        w.enable_timing(True)
        w.kernel_time_ms()
        l0 = w.launch_count()
        clk.begin()
        ms = timed(step, args.steps, 0)
        clk.end()
    launches = w.launch_count() - l0
    das_ms, das_n, pack_ms, pack_n = w.kernel_time_ms()
    w.enable_timing(False)
    value = B * args.steps / (ms / 1e3)
    kinfo = w.kernel_info()

    # ---- end to end: host buffers through the C ABI (N = 1) / pinned copies + gather (N > 1) ----
    e2e = None
    if not args.no_e2e:
        if world == 1:
            def e2e_step():
                w.power_map_batch_ptr(host_in.data_ptr(), T, B, host_out.data_ptr())
        else:
            # The rank's frame slice goes up in K chunks of frames: chunk j+1 is uploaded (C / G_d channel rows per rank of
            # the frame group, over its own PCIe link) and replicated inside the group by an all-gather over NVLink on a
            # copy stream while chunk j is being computed -- what bflk_power_map_batch does inside one process.
            assert C % gd == 0
            rows = C // gd
            K = max(1, min(4, nf // 64))
            per_chunk = -(-nf // K)
            per_chunk += per_chunk & 1
            chunks = [(a, min(per_chunk, nf - a)) for a in range(0, nf, per_chunk)]
            host_np = host_in.numpy()
            host_chunks, slice_devs, chunk_devs, chunk_T = [], [], [], []
            for a, n in chunks:
                Tj = (n - 1) * N + c["W"]
                part = np.ascontiguousarray(host_np[dgrp * rows:(dgrp + 1) * rows, a * N: a * N + Tj])
                host_chunks.append(torch.from_numpy(part).pin_memory())
                slice_devs.append(torch.empty((rows, Tj), dtype=torch.float32, device=dev))
                chunk_devs.append(slice_devs[-1] if gd == 1 else torch.empty((C, Tj), dtype=torch.float32, device=dev))
                chunk_T.append(Tj)
            copy_stream = torch.cuda.Stream(device=dev)
            events = [torch.cuda.Event() for _ in chunks]
            row_bytes = local_tight.stride(0) * 4

            def e2e_step():
                with torch.cuda.stream(copy_stream):
                    for j in range(len(chunks)):
                        slice_devs[j].copy_(host_chunks[j], non_blocking=True)
                        if gd > 1:
                            dist.all_gather_into_tensor(chunk_devs[j].view(gd * rows, chunk_T[j]), slice_devs[j], group=in_group)
                        events[j].record(copy_stream)
                for j, (a, n) in enumerate(chunks):
                    work_stream.wait_event(events[j])
                    w.power_map_batch_dev(chunk_devs[j].data_ptr(), chunk_T[j], n, local_tight.data_ptr() + a * row_bytes, cs)
                if not tight:
                    local_pow[:nf, :count].copy_(local_tight)
                dist.all_gather_into_tensor(gathered.view(world * nf_max, per), local_pow)
                host_out.copy_(shard.assemble_2d(gathered, B, D, gd, gf), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                copy_stream.synchronize()
        e2e_steps = max(3, args.steps // 3)
        if world == 1:
            # synchronous host API (returns after the D2H copy): host wall clock brackets the whole call
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            ems = 1e3 * (time.perf_counter() - t0)
        else:
            ems = timed(e2e_step, e2e_steps, 1)
        e2e = {"value": B * e2e_steps / (ems / 1e3), "unit": UNIT, "h2d_bytes_per_step": C * T_loc * 4 * gf,
               "d2h_bytes_per_step": B * D * 4, "steps": e2e_steps,
               "path": "bflk_power_map_batch (host buffers)" if world == 1 else
               "per frame chunk: pinned H2D of C/G_d channel rows per rank + all_gather inside the frame group (copy stream) "
               "overlapping bflk_power_map_batch_dev of the previous chunk; then map all_gather + D2H; h2d bytes summed over the ranks"}
        # outside the timed region: the maps the end-to-end path delivered are the ones the device-resident path computes
        e2e_maps = host_out.clone()
        stream_dev.copy_(host_in)
        step()
        torch.cuda.current_stream().synchronize()
        ref_maps = (local_tight if world == 1 else shard.assemble_2d(gathered, B, D, gd, gf)).cpu()
        e2e["same_maps_as_resident_path"] = bool(torch.equal(e2e_maps, ref_maps))

    if rank == 0:
        pk, pk_kind = peaks()
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        clocks = clk.summary()
        max_mhz = pk.get("sm_max_mhz") or clocks["sm_max_mhz"] or 1965.0
        peak_tf = sm_count * FP32_LANES_PER_SM * 2 * max_mhz * 1e6 / 1e12
        # algorithmic FLOPs one das launch performs on this rank: B maps x per directions
        fl = flops_per_map(C, count, N) * nf
        das_avg_s = das_ms / 1e3 / max(1, das_n)
        achieved_tf = fl / das_avg_s / 1e12 if das_n else None
        alg_bytes = 4 * C * T_loc + 4 * nf * count
        roof = {"bound": "fp32", "kernel": {1: "das_generic", 2: "das_tile", 3: "das_bcast", 4: "das_tile_fma2"}.get(kinfo[0], "?"), "achieved": achieved_tf,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if achieved_tf else None,
                "peak_source": f"{sm_count} SMs x 128 FP32 lanes x 2 x sm_max_mhz {max_mhz:.0f} ({pk_kind} MEASURED_PEAKS.json clock)",
                "frac_at_sampled_clock": (achieved_tf / (peak_tf * clocks["sm_mhz"] / max_mhz)) if achieved_tf and clocks["sm_mhz"] else None,
                "flop_per_launch": fl, "avg_launch_ms": das_avg_s * 1e3, "launches_timed": das_n,
                "kernel_share_of_step": das_ms / ms, "pack_share_of_step": pack_ms / ms, "traffic": None}
        try:        # DRAM bytes per launch from the committed ncu capture of this exact command, if there is one
            t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(name)
            if t and t["frames_per_step"] == B and t["kernel"] == roof["kernel"] and t["n_gpus"] == world:
                roof["traffic"] = t["dram_bytes_per_launch"]
                roof["traffic_source"] = t["source"]
        except Exception:
            pass
        roof_hbm = {"bound": "hbm", "achieved": alg_bytes / das_avg_s / 1e9 if das_n else None, "peak": pk["hbm_gbs"],
                    "unit": "GB/s", "frac": alg_bytes / das_avg_s / 1e9 / pk["hbm_gbs"] if das_n else None,
                    "bytes_per_launch": alg_bytes, "peak_source": pk_kind}
        cfg.update(parallelism=f"{gd} direction groups x {gf} frame groups" if world > 1 else "single GPU", directions_per_gpu=count,
                   frames_per_gpu=nf,
                   kernel=roof["kernel"], tile_span=kinfo[1], window_chunks=kinfo[2])
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "roofline": roof, "roofline_hbm": roof_hbm,
                "dir_samples_per_sec": value * D * N, "tflops_whole_job": value * flops_per_map(C, D, N) / 1e12}
        if world == 1 and not args.no_cpu_baseline:
            try:
                n_threads = os.cpu_count() or 1
                rate, sample = cpu_reference_rate(c, host_in.numpy()[:, :16 * N + c["W"]], n_threads, budget_s=3.0)
                rate1, sample1 = cpu_reference_rate(c, host_in.numpy()[:, :16 * N + c["W"]], 1, budget_s=2.0)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": n_threads, "kind": "reference",
                                        "sample": sample, "single_thread_value": rate1, "single_thread_sample": sample1}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    c = cases.CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, c, args.config)
    else:
        run_ours(args, c, args.config)


if __name__ == "__main__":
    main()
