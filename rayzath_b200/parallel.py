"""One process per GPU: how a frame is sharded over the GPUs of one box and how the shards are combined.

The path shards naturally: pixels and samples never interact; the only shared state is the
read-only scene and the per-pixel accumulator, whose rgb sum and alpha (completed-path count) are both additive
(cuda_render_kernel.cu:100-105). So there is NO data-path collective while rendering:
  * sample split: every rank renders the same pixels with its own RNG stream (`stream_seed`);
  * tile split:   contiguous row bands per rank (`row_band`), for 4K frames;
and ONE exchange step at resolve time: sum the float4 accumulators onto the root and tone-map there. Two
implementations of that step:
  * `reduce_accum`        torch.distributed.reduce (NCCL over NVLink; gloo on CPU tensors in the tests)
  * `FusedResolve`        the B200-native one: ranks export their accumulator as CUDA IPC handles, the root maps
                          them and ONE kernel (k_tonemap) loads the peers' pixels over NVLink, sums and tone-maps.
The reference is single-GPU (device 0, cuda_engine_core.cu:17); this module has no counterpart there.
"""
from __future__ import annotations

import os
from typing import List, Tuple

GOLDEN64 = 0x9E3779B97F4A7C15


def env_ranks() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group(backend=backend)
    return dist


def stream_seed(base_seed: int, rank: int) -> int:
    """Disjoint RNG streams: the device RNG hashes (seed, pixel, pass, dimension), so distinct seeds are distinct
    streams; ranks are spread with the 64-bit golden ratio so that nearby base seeds of different jobs do not collide."""
    return (int(base_seed) + int(rank) * GOLDEN64) & 0xFFFFFFFFFFFFFFFF


def passes_for_rank(total_passes: int, world: int, rank: int) -> int:
    """Strong-scaling split of a pass budget (remainder to the low ranks)."""
    return total_passes // world + (1 if rank < total_passes % world else 0)


def row_band(height: int, world: int, rank: int) -> Tuple[int, int]:
    """Tile split: [row_begin, row_end) of this rank; bands are contiguous, disjoint and cover the frame."""
    base, rem = divmod(height, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_accum(accum, dst: int = 0):
    """Sum the [h, w, 4] float32 accumulators of all ranks onto `dst` (in place there)."""
    import torch.distributed as dist
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


class FusedResolve:
    """The exchange step as ONE kernel over NVLink peer memory, stream-ordered (no host synchronisation between the
    ranks inside the step):

      setup (once per frame-buffer allocation): every rank exports its accumulator as a CUDA IPC handle; handles are
      exchanged through the process group; the root maps them lazily on first use (rzb_resolve_ipc caches them).
      call: [render passes ... on the stream] -> token all_reduce (an NCCL barrier in stream order: it completes on
      the root only after every rank's preceding work has completed) -> root: k_tonemap loads the peers' float4
      pixels over NVLink, sums them with its own and tone-maps, RGBA8 + depth to HOST buffers -> token all_reduce
      (peers may not touch their accumulators before the root has read them).
    All ranks must call it. Returns (rgba8, depth) on the root, None elsewhere."""

    def __init__(self, ctx):
        import torch
        import torch.distributed as dist
        self.ctx, self.dist = ctx, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        handles: List[bytes] = [b""] * self.world
        dist.all_gather_object(handles, ctx.accum_ipc_handle())
        self.peer_handles = [h for r, h in enumerate(handles) if r != 0]
        self.token = torch.zeros(1, dtype=torch.float32, device="cuda:%d" % ctx.device)

    def __call__(self, want_depth: bool = False):
        self.dist.all_reduce(self.token)
        out = None
        if self.rank == 0:
            out = self.ctx.resolve_ipc(self.peer_handles, want_depth=want_depth)
        self.dist.all_reduce(self.token)
        return out
