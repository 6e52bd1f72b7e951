"""Direction sharding of the steering grid across the GPUs of one box (one process per GPU).

Each direction's power is independent (src/dsp/mimo.cpp:121-151), so rank g of G owns a contiguous run of
the row-major grid and the per-rank slices are assembled with one all-gather per batch.  Works with any
torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def direction_shard(n_directions, world, rank):
    """(first, count) of rank's contiguous run; the first n_directions % world ranks get one extra."""
    base, extra = divmod(n_directions, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def shard_counts(n_directions, world):
    return [direction_shard(n_directions, world, r)[1] for r in range(world)]


def padded_count(n_directions, world):
    return -(-n_directions // world)


def gather_maps(local, n_directions, out=None, group=None):
    """local: [B][padded_count] (this rank's slice, zero-padded to the common width) -> [world][B][padded]."""
    world = dist.get_world_size(group)
    B, padded = local.shape
    if out is None:
        out = torch.empty((world, B, padded), dtype=local.dtype, device=local.device)
    # concatenation along dim 0 is the one output form every backend accepts
    dist.all_gather_into_tensor(out.view(world * B, padded), local.contiguous(), group=group)
    return out


def assemble(gathered, n_directions):
    """[world][B][padded] -> [B][n_directions] in grid order (drops the padding of ragged shards)."""
    world, B, padded = gathered.shape
    counts = shard_counts(n_directions, world)
    if all(c == padded for c in counts):
        return gathered.permute(1, 0, 2).reshape(B, world * padded)
    return torch.cat([gathered[r, :, :counts[r]] for r in range(world)], dim=1)


def channel_slice(n_channels, world, rank):
    """(first, count) of the channel rows rank uploads in replicate_input (equal slices; the last rank takes the rest)."""
    per = -(-n_channels // world)
    first = min(rank * per, n_channels)
    return first, min(per, n_channels - first)


def replicate_input(host_stream, dev_stream, group=None, staging=None):
    """Every rank needs ALL channels of the batch (the grid is sharded by direction, not by channel).  Instead of
    each rank pulling the whole [C][T] stream over its own PCIe link, rank g uploads only its C / G channel rows
    (G links in parallel) and one all-gather over NVLink replicates them: the channel-major layout makes the
    concatenation of the slices the stream itself.  host_stream: pinned [C][T] tensor, dev_stream: [C][T] on the
    device.  Falls back to a plain full copy when C is not a multiple of the world size."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    C, T = host_stream.shape
    if world == 1 or C % world:
        dev_stream.copy_(host_stream, non_blocking=True)
        return dev_stream
    first, count = channel_slice(C, world, rank)
    if staging is None:
        staging = torch.empty((count, T), dtype=dev_stream.dtype, device=dev_stream.device)
    staging.copy_(host_stream[first:first + count], non_blocking=True)
    dist.all_gather_into_tensor(dev_stream.view(world * count, T), staging, group=group)
    return dev_stream


# ---- batches: direction groups x frame groups ----------------------------------------------------------------------
# A batch of B frames has a second independent axis.  Sharding ONLY the grid makes every rank repeat the per-batch
# pre-pass (the pack of the whole input: 3 % of a step on one GPU, 20 % on eight); sharding only the frames would leave
# the grid whole.  The bench therefore arranges G ranks as G_d direction groups x G_f frame groups: rank r works on
# direction slice r % G_d of frame slice r // G_d, the pre-pass is repeated G_d times instead of G times, and one
# all-gather over all ranks assembles [B][D].  G_d = G is the pure grid sharding above.
def grid_2d(world, dir_groups=0):
    """(G_d, G_f): dir_groups if given (must divide world), else 2 direction groups when world is even."""
    gd = dir_groups if dir_groups else (2 if world % 2 == 0 else 1)
    if gd < 1 or world % gd:
        raise ValueError(f"{gd} direction groups do not divide {world} ranks")
    return gd, world // gd


def frame_shard(n_frames, groups, g):
    """(first, count) of frame group g: contiguous, even counts (the kernel works on block pairs), the last takes the rest."""
    per = -(-n_frames // groups)
    per += per & 1
    first = min(g * per, n_frames)
    return first, max(0, min(per, n_frames - first))


def assemble_2d(gathered, n_frames, n_directions, gd, gf):
    """[world][frames per group][padded directions] (rank r = frame group r // gd, direction group r % gd) -> [B][D]."""
    world, nf, padded = gathered.shape
    assert world == gd * gf
    counts = shard_counts(n_directions, gd)
    fcounts = [frame_shard(n_frames, gf, g)[1] for g in range(gf)]
    if all(c == padded for c in counts) and all(c == nf for c in fcounts):
        return gathered.view(gf, gd, nf, padded).permute(0, 2, 1, 3).reshape(gf * nf, gd * padded)
    rows = [torch.cat([gathered[g * gd + d, :fcounts[g], :counts[d]] for d in range(gd)], dim=1) for g in range(gf)]
    return torch.cat(rows, dim=0)
