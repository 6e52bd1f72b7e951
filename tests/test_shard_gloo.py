"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo processes shard the steering grid, each computes its
slice (the CPU oracle stands in for the kernel), slices are all-gathered and assembled into the full map."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rows, cols, B, out_path):
    import sys
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bflk import shard, synth
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz = O.create_antenna()
    off, fr = O.mimo_lut(xyz, rows, cols, 180.0)
    D = rows * cols
    stream = synth.make_stream(synth.tile_geometry(cases.origins(1, 1)), (B - 1) * 256 + 1024)
    first, count = shard.direction_shard(D, world, rank)
    padded = shard.padded_count(D, world)
    local = torch.zeros((B, padded), dtype=torch.float32)
    for b in range(B):
        w = np.ascontiguousarray(stream[:, b * 256: b * 256 + 1024])
        local[b, :count] = torch.from_numpy(O.mimo_update(w, off[first:first + count], fr[first:first + count]))
    full = shard.assemble(shard.gather_maps(local, D), D)
    if rank == 0:
        ref = np.stack([O.mimo_update(np.ascontiguousarray(stream[:, b * 256: b * 256 + 1024]), off, fr) for b in range(B)])
        np.save(out_path, np.stack([full.numpy(), ref]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,rows,cols", [(2, 8, 8), (3, 5, 7)])
def test_sharded_maps_assemble_to_full_map(tmp_path, world, rows, cols):
    out = str(tmp_path / "maps.npy")
    mp.spawn(_worker, args=(world, _free_port(), rows, cols, 3, out), nprocs=world, join=True)
    full, ref = np.load(out)
    assert full.shape == (3, rows * cols)
    assert np.array_equal(full, ref)          # same oracle arithmetic per direction -> bit-identical after assembly


def _input_worker(rank, world, port, C, T, out_path):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bflk import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host = torch.arange(C * T, dtype=torch.float32).reshape(C, T) * 0.5
    devt = torch.full((C, T), -1.0)
    shard.replicate_input(host, devt)
    ok = torch.equal(devt, host)
    flags = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([float(ok)]))
    if rank == 0:
        np.save(out_path, np.array([f.item() for f in flags]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,C", [(2, 64), (4, 64), (3, 64)])      # 64 % 3 != 0: full-copy fallback
def test_channel_sliced_upload_replicates_the_stream(tmp_path, world, C):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_input_worker, args=(world, _free_port(), C, 300, out), nprocs=world, join=True)
    assert np.all(np.load(out) == 1.0)


def _grid2d_worker(rank, world, port, gd, B, D, out_path):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bflk import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gd, gf = shard.grid_2d(world, gd)
    d0, dc = shard.direction_shard(D, gd, rank % gd)
    f0, fc = shard.frame_shard(B, gf, rank // gd)
    per, nf = shard.padded_count(D, gd), shard.frame_shard(B, gf, 0)[1]
    truth = torch.arange(B * D, dtype=torch.float32).reshape(B, D)          # "map" value = its own flat index
    local = torch.zeros((nf, per))
    local[:fc, :dc] = truth[f0:f0 + fc, d0:d0 + dc]
    out = torch.empty((world, nf, per))
    dist.all_gather_into_tensor(out.view(world * nf, per), local)
    full = shard.assemble_2d(out, B, D, gd, gf)
    ok = torch.equal(full, truth)
    flags = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([float(ok)]))
    if rank == 0:
        np.save(out_path, np.array([f.item() for f in flags]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,gd,B,D", [(4, 2, 8, 64), (4, 2, 10, 35), (2, 1, 6, 16), (3, 3, 4, 10)])
def test_direction_by_frame_grid_assembles(tmp_path, world, gd, B, D):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_grid2d_worker, args=(world, _free_port(), gd, B, D, out), nprocs=world, join=True)
    assert np.all(np.load(out) == 1.0)


def test_direction_shard_partition():
    import sys
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from bflk import shard
    for D in (1, 7, 1024, 1025, 65536):
        for world in (1, 2, 3, 4, 8):
            runs = [shard.direction_shard(D, world, r) for r in range(world)]
            assert runs[0][0] == 0 and sum(c for _, c in runs) == D
            for (f0, c0), (f1, _) in zip(runs, runs[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in runs) == shard.padded_count(D, world)
            assert max(c for _, c in runs) - min(c for _, c in runs) <= 1


def test_frame_groups_and_bench_batch_rule():
    """Host logic of the direction x frame decomposition and of bench.py's default batch: frame slices are even-sized,
    contiguous and cover the batch; the default batch makes the block pairs a multiple of 148 SMs (whole CTA waves at
    1, 2, 4 and 8 GPUs) and is larger than L2."""
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bflk import shard
    import bench
    assert shard.grid_2d(1) == (1, 1) and shard.grid_2d(2) == (2, 1) and shard.grid_2d(8) == (2, 4)
    assert shard.grid_2d(8, 8) == (8, 1) and shard.grid_2d(8, 1) == (1, 8) and shard.grid_2d(3) == (1, 3)
    with pytest.raises(ValueError):
        shard.grid_2d(8, 3)
    for B in (2, 7, 592, 593, 4144):
        for groups in (1, 2, 3, 4, 8):
            runs = [shard.frame_shard(B, groups, g) for g in range(groups)]
            assert runs[0][0] == 0 and sum(c for _, c in runs) == B
            for (f0, c0), (f1, c1) in zip(runs, runs[1:]):
                assert f0 + c0 == f1 or c1 == 0
            nonempty = [c for _, c in runs if c]
            assert all(c % 2 == 0 for c in nonempty[:-1]) and runs[0][1] == max(nonempty)   # only the last slice may be odd
    for name in ("cfg1", "cfg2", "cfg3"):
        c = cases.CONFIGS[name]
        B = bench.default_frames(c)
        assert (B // 2) % 148 == 0 and B % 2 == 0
        assert 64 * c["nx"] * c["ny"] * B * c["N"] * 4 >= 256e6          # input larger than L2 (126 MB)
    assert bench.default_frames(cases.CONFIGS["cfg3"]) == 592
    for world in (2, 4, 8):                                               # every rank gets whole waves at cfg3
        gd, gf = shard.grid_2d(world)
        nf = shard.frame_shard(592, gf, 0)[1]
        ctas = (nf // 2) * -(-(1024 // gd // 4) // 16)
        assert ctas % 148 == 0


# ---- the plan the LIBRARY uses (bflk_shard_plan through the C ABI; pure host arithmetic, no GPU) -------------------------
def test_c_abi_shard_plan_matches_host_logic():
    import bflk
    from bflk import shard
    for D in (1, 7, 35, 1024, 1025, 65536):
        for B in (1, 2, 7, 592, 593):
            for world in (1, 2, 3, 4, 8):
                for gd in [0] + [g for g in (1, 2, 3, 4, 8) if world % g == 0]:
                    egd, egf = shard.grid_2d(world, gd)
                    cover = np.zeros((B, D), np.int32)
                    for r in range(world):
                        d0, dc, f0, fc = bflk.shard_plan(D, B, world, r, gd)
                        assert (d0, dc) == shard.direction_shard(D, egd, r % egd)
                        assert (f0, fc) == shard.frame_shard(B, egf, r // egd)
                        cover[f0:f0 + fc, d0:d0 + dc] += 1
                    assert np.all(cover == 1)                       # every (frame, direction) computed exactly once
    with pytest.raises(bflk.BflkError):
        bflk.shard_plan(1024, 16, 8, 0, 3)                          # 3 direction groups do not divide 8 ranks
    with pytest.raises(bflk.BflkError):
        bflk.shard_plan(1024, 16, 4, 4, 0)                          # rank out of range


def _abi_plan_worker(rank, world, port, gd, rows, cols, B, out_path):
    """World-size-N job whose per-rank slice comes from the C ABI (bflk_shard_plan); the CPU oracle stands in for the
    kernel, gloo for NCCL; the gathered slices are assembled exactly like multi.cu's assemble_kernel does."""
    import sys
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import bflk
    from bflk import synth
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz = O.create_antenna()
    off, fr = O.mimo_lut(xyz, rows, cols, 180.0)
    D = rows * cols
    stream = synth.make_stream(synth.tile_geometry(cases.origins(1, 1)), (B - 1) * 256 + 1024)
    plans = [bflk.shard_plan(D, B, world, r, gd) for r in range(world)]
    d0, dc, f0, fc = plans[rank]
    per, nf = max(p[1] for p in plans), max(p[3] for p in plans)
    local = torch.zeros((nf, per), dtype=torch.float32)
    for b in range(fc):
        w = np.ascontiguousarray(stream[:, (f0 + b) * 256: (f0 + b) * 256 + 1024])
        local[b, :dc] = torch.from_numpy(O.mimo_update(w, off[d0:d0 + dc], fr[d0:d0 + dc]))
    gathered = torch.empty((world, nf, per), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered.view(world * nf, per), local)
    full = torch.zeros((B, D))
    for r, (rd0, rdc, rf0, rfc) in enumerate(plans):
        full[rf0:rf0 + rfc, rd0:rd0 + rdc] = gathered[r, :rfc, :rdc]
    if rank == 0:
        ref = np.stack([O.mimo_update(np.ascontiguousarray(stream[:, b * 256: b * 256 + 1024]), off, fr) for b in range(B)])
        np.save(out_path, np.stack([full.numpy(), ref]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,gd,rows,cols,B", [(2, 0, 6, 6, 4), (2, 1, 5, 7, 3), (4, 2, 4, 5, 5)])
def test_world_size_n_job_through_the_c_abi_plan(tmp_path, world, gd, rows, cols, B):
    out = str(tmp_path / "maps.npy")
    mp.spawn(_abi_plan_worker, args=(world, _free_port(), gd, rows, cols, B, out), nprocs=world, join=True)
    full, ref = np.load(out)
    assert np.array_equal(full, ref)
