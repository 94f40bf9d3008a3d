"""Host-side logic of the N>1 path on CPU (gloo, world_size 2): stream seeds, pass/tile split, and the one
exchange step -- the accumulator reduce followed by the tone map on the root -- checked against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

from rayzath_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_seeds_are_distinct():
    seeds = {parallel.stream_seed(b, r) for b in range(16) for r in range(8)}
    assert len(seeds) == 16 * 8
    assert parallel.stream_seed(7, 0) == 7


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_pass_and_tile_split_cover_exactly(world):
    for total in (0, 1, 7, 256, 4096):
        assert sum(parallel.passes_for_rank(total, world, r) for r in range(world)) == total
    for height in (1, 7, 1080, 2160):
        bands = [parallel.row_band(height, world, r) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == height
        assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in bands) - min(e - b for b, e in bands) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    from rayzath_b200 import parallel as par
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(par.stream_seed(99, rank) % (2 ** 32))
    acc = rng.random((9, 16, 4)).astype(np.float32) * 5
    acc[..., 3] = np.floor(acc[..., 3]) + 1
    np.save(os.path.join(out_dir, "acc_%d.npy" % rank), acc)
    t = torch.from_numpy(acc.copy())
    par.reduce_accum(t, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_then_tonemap_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    import rz_oracle as O
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(str(tmp_path / ("acc_%d.npy" % r))) for r in range(world)]
    total = np.load(str(tmp_path / "sum.npy"))
    assert np.array_equal(total, parts[0] + parts[1])
    assert not np.array_equal(parts[0], parts[1])
    # the resolve on the root tone-maps the SUM; alpha (path count) adds up like the colour sums
    img = O.tonemap(total, 0.01, 0.006)
    assert img.shape == (9, 16, 4) and (img[..., 3] == 255).all()
    assert np.array_equal(img, O.tonemap(parts[0] + parts[1], 0.01, 0.006))
