"""One process per rank, sample streams sharded over the ranks, accumulators combined by the fused IPC resolve
(rank 0 loads the peers' pixels through CUDA IPC mappings -- NVLink between GPUs -- sums and tone-maps in one kernel)
and by the sliced resolve (every rank sums and tone-maps its slice of the frame; flag barriers in peer memory, no NCCL).
Runs with 2 ranks on 2 GPUs (NCCL) when the box has them, else with 2 ranks sharing GPU 0 (gloo for the host-side
ordering; the IPC path is the same). Checked bit-exactly against the oracle's tone map of the summed accumulators."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_gpus, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from rayzath_b200 import capi, parallel, scenes
    dev = rank % n_gpus
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if n_gpus >= world else "gloo", rank=rank, world_size=world)
    w = scenes.materials_scene(resolution=(160, 90), res=16)
    ctx = capi.Context(dev)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_scene(w.flatten())
    ctx.set_camera(w.camera_struct())
    ctx.set_config(1, 1, 8, capi.FLAG_NONE, parallel.stream_seed(77, rank))
    ctx.reset()
    ctx.render(parallel.passes_for_rank(25, world, rank))
    fused = parallel.FusedResolve(ctx)
    out = fused(want_depth=True)
    np.save(os.path.join(out_dir, "acc_%d.npy" % rank), ctx.read_accum())
    # the sliced resolve (one kernel per rank: flag barrier in peer memory, slice sum over NVLink, tone map into rank
    # 0's staging image), twice: the flags carry the call number
    rgba_pin = torch.empty((90, 160, 4), dtype=torch.uint8, pin_memory=True).numpy()
    depth_pin = torch.empty((90, 160), dtype=torch.float32, pin_memory=True).numpy()
    sliced = parallel.SlicedResolve(ctx, rgba_pin, depth_pin)
    for _ in range(2):
        rgba_pin[:] = 0
        sliced()
        sliced.wait()
    assert sliced.last_ms() is not None and sliced.last_ms() >= 0.0
    if rank == 0:
        np.save(os.path.join(out_dir, "rgba_sliced.npy"), rgba_pin.copy())
        np.save(os.path.join(out_dir, "depth_sliced.npy"), depth_pin.copy())
        np.save(os.path.join(out_dir, "depth_fused.npy"), out[1])
    if rank == 0:
        np.save(os.path.join(out_dir, "rgba.npy"), out[0])
        # the NCCL/gloo reduce path must give the same sum
        total = ctx.accum_tensor().clone()
    else:
        total = ctx.accum_tensor().clone()
    parallel.reduce_accum(total, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), total.cpu().numpy())
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_fused_ipc_resolve_two_ranks(tmp_path):
    import torch
    import torch.multiprocessing as mp
    import rz_oracle as O
    from rayzath_b200 import scenes
    n_gpus = torch.cuda.device_count()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_gpus, str(tmp_path)), nprocs=world, join=True)
    a0, a1 = np.load(tmp_path / "acc_0.npy"), np.load(tmp_path / "acc_1.npy")
    assert not np.array_equal(a0, a1)
    assert a0[..., 3].sum() > 0 and a1[..., 3].sum() > 0
    cam = scenes.materials_scene(resolution=(160, 90), res=16).camera_struct()[0]
    expect = O.tonemap(a0 + a1, float(cam["aperture"]), float(cam["exposure_time"]))
    assert np.array_equal(np.load(tmp_path / "rgba.npy"), expect)
    assert np.array_equal(np.load(tmp_path / "rgba_sliced.npy"), expect)
    assert np.array_equal(np.load(tmp_path / "depth_sliced.npy"), np.load(tmp_path / "depth_fused.npy"))
    assert np.array_equal(np.load(tmp_path / "reduced.npy"), a0 + a1)
