"""Multi-GPU through the C ABI (libbflk.so + NCCL): needs >= 2 B200s in one box (gpurun --gpus 2); skipped otherwise.
A single-process group (bflk_group_*) shards the steering grid x the frames of a batch over the devices, all-gathers the
slices with NCCL and must reproduce the single-GPU maps bit for bit (same kernels, same per-direction arithmetic)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs in one box")


def _stream(c, B):
    from bflk import synth
    xyz = synth.tile_geometry(cases.origins(c["nx"], c["ny"]))
    T = (B - 1) * c["N"] + c["W"]
    return synth.make_stream(xyz, T + (T & 1))


@needs2
@pytest.mark.parametrize("dir_groups", [0, 1, 2])
@pytest.mark.parametrize("name,rows,cols,B", [("cfg2", 16, 12, 7), ("cfg1", 9, 9, 4), ("cfg3", 32, 32, 10)])
def test_group_host_batch_equals_single_gpu(name, rows, cols, B, dir_groups):
    import bflk
    c = cases.CONFIGS[name]
    stream = _stream(c, B)
    single = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), rows, cols, c["fov"])
    ref = single.power_map_batch(stream, B)
    n = min(_n_gpus(), 2)
    g = bflk.Group(cases.origins(c["nx"], c["ny"]), rows, cols, c["fov"], devices=list(range(n)), dir_groups=dir_groups)
    assert g.size() == n
    got = g.power_map_batch(stream, B)
    assert np.array_equal(got, ref)
    got2 = g.power_map_batch(stream, B)                 # second call: buffers and plan reused
    assert np.array_equal(got2, ref)
    g.close()


@needs2
def test_group_device_resident_all_ranks_get_full_maps():
    import torch
    import bflk
    c = cases.CONFIGS["cfg3"]
    B = 12
    stream = _stream(c, B)
    single = bflk.MIMOWorker(cases.origins(c["nx"], c["ny"]), c["rows"], c["cols"], c["fov"])
    ref = single.power_map_batch(stream, B)
    n = _n_gpus()
    for gd in sorted({1, 2, n}):
        if n % gd:
            continue
        g = bflk.Group(cases.origins(c["nx"], c["ny"]), c["rows"], c["cols"], c["fov"], devices=list(range(n)), dir_groups=gd)
        ins = [torch.from_numpy(stream).to(f"cuda:{i}") for i in range(n)]
        outs = [torch.zeros((B, c["rows"] * c["cols"]), dtype=torch.float32, device=f"cuda:{i}") for i in range(n)]
        g.power_map_batch_dev([t.data_ptr() for t in ins], stream.shape[1], B, [t.data_ptr() for t in outs])
        for o in outs:
            assert np.array_equal(o.cpu().numpy(), ref)
        g.close()


@needs2
def test_pipelined_device_batches_one_process_per_gpu():
    """bflk_comm_init_rank (one process per GPU, as torchrun forms them) + bflk_power_map_batch_sharded_dev_submit / _join:
    the check itself is tests/multi_dev_pipeline_check.py, run here under torch.distributed.run on two GPUs."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tests", "multi_dev_pipeline_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PIPELINE_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
