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
