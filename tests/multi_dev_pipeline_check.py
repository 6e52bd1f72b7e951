"""Run under torchrun (one process per GPU) by tests/test_gpu_multi.py: bflk_power_map_batch_sharded_dev_submit / _join --
batches whose kernels run under the previous batch's all-gather -- must deliver the bits of the synchronous call, for
alternating inputs, several batches in flight, a synchronous call in between and a direction split or a frame split."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "beamforming-lk_b200"), os.path.join(ROOT, "tests")]
import bflk  # noqa: E402
from bflk import synth  # noqa: E402
import cases  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("gloo")
c = cases.CONFIGS["cfg2"]
rows, cols, B = 24, 20, 10
T = (B - 1) * c["N"] + c["W"]
xyz = synth.tile_geometry(cases.origins(c["nx"], c["ny"]))
base = synth.make_stream(xyz, T)
inputs = [torch.from_numpy(base * np.float32(1.0 + 0.5 * k)).to(dev) for k in range(3)]
for dir_groups in (0, 1):
    w = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), rows, cols, c["fov"], device=local)
    box = [bflk.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    w.comm_init_rank(box[0], world, rank, dir_groups)
    stream = torch.cuda.Stream(device=dev)
    cs = stream.cuda_stream
    want = []
    for x in inputs:
        o = torch.zeros((B, rows * cols), dtype=torch.float32, device=dev)
        w.power_map_batch_sharded_dev(x.data_ptr(), T, B, o.data_ptr(), cs)
        stream.synchronize()
        want.append(o.cpu().numpy())
    assert not np.array_equal(want[0], want[1])
    outs = [torch.zeros((B, rows * cols), dtype=torch.float32, device=dev) for _ in range(7)]
    order = [0, 1, 2, 1, 0, 2, 2]
    for k, i in enumerate(order):                       # seven batches back to back, two buffer sets
        w.power_map_batch_sharded_dev_submit(inputs[i].data_ptr(), T, B, outs[k].data_ptr(), cs)
    w.power_map_batch_sharded_dev_join(cs)
    stream.synchronize()
    for k, i in enumerate(order):
        assert np.array_equal(outs[k].cpu().numpy(), want[i]), (dir_groups, k)
    # a synchronous call between submitted batches, then more submits into one output buffer (ordered on the comm stream)
    w.power_map_batch_sharded_dev_submit(inputs[0].data_ptr(), T, B, outs[0].data_ptr(), cs)
    w.power_map_batch_sharded_dev(inputs[1].data_ptr(), T, B, outs[1].data_ptr(), cs)
    w.power_map_batch_sharded_dev_submit(inputs[2].data_ptr(), T, B, outs[2].data_ptr(), cs)
    w.power_map_batch_sharded_dev_submit(inputs[0].data_ptr(), T, B, outs[2].data_ptr(), cs)
    w.power_map_batch_sharded_dev_join(cs)
    stream.synchronize()
    assert np.array_equal(outs[0].cpu().numpy(), want[0]) and np.array_equal(outs[1].cpu().numpy(), want[1])
    assert np.array_equal(outs[2].cpu().numpy(), want[0])
    # and the host-batch path afterwards still agrees
    host_in = torch.from_numpy(base).pin_memory()
    host_out = torch.zeros((B, rows * cols), dtype=torch.float32).pin_memory()
    w.power_map_batch_sharded_dev_submit(inputs[1].data_ptr(), T, B, outs[3].data_ptr(), cs)      # still in flight when ...
    w.power_map_batch_sharded_ptr(host_in.data_ptr(), T, B, host_out.data_ptr())                  # ... the host path runs
    assert np.array_equal(host_out.numpy(), want[0])
    w.power_map_batch_sharded_dev_join(cs)
    stream.synchronize()
    assert np.array_equal(outs[3].cpu().numpy(), want[1])
    dist.barrier()
    w.close()
if rank == 0:
    print("PIPELINE_OK")
dist.destroy_process_group()
