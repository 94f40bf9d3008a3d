"""One process per GPU: how a frame is sharded over the GPUs of one box and how the shards are combined.

The path shards naturally: pixels and samples never interact; the only shared state is the
read-only scene and the per-pixel accumulator, whose rgb sum and alpha (completed-path count) are both additive
(cuda_render_kernel.cu:100-105). So there is NO data-path collective while rendering:
  * sample split: every rank renders the same pixels with its own RNG stream (`stream_seed`);
  * tile split:   contiguous row bands per rank (`row_band`), for 4K frames;
and ONE exchange step at resolve time: sum the float4 accumulators onto the root and tone-map there. Two
implementations of that step:
  * `reduce_accum`        torch.distributed.reduce (NCCL over NVLink; gloo on CPU tensors in the tests)
  * `SlicedResolve`       the B200-native one (bench.py's default): ranks export their accumulator and a small
                          exchange buffer as CUDA IPC handles; per resolve ONE kernel per rank does a flag barrier in
                          peer memory, sums ITS slice of all accumulators over NVLink (all-to-all ingress), tone-maps it
                          into rank 0's staging image and does the closing barrier -- no NCCL call, no host round trip
  * `FusedResolve`        its predecessor: the root alone pulls all peers' pixels (incast on one GPU's NVLink ingress),
                          ordered by two NCCL token all-reduces
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
    All ranks must call it. Returns (rgba8, depth) on the root, None elsewhere.

    Precondition: the context renders on torch's CURRENT stream (ctx.set_stream(torch.cuda.current_stream().cuda_stream)),
    because the token all-reduce orders only work of the stream NCCL runs on; a context on its private stream is
    synchronised with the host before the first token instead (correct, but the host waits)."""

    def __init__(self, ctx):
        import torch
        import torch.distributed as dist
        self.ctx, self.dist = ctx, dist
        self._torch = torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        handles: List[bytes] = [b""] * self.world
        dist.all_gather_object(handles, ctx.accum_ipc_handle())
        self.peer_handles = [h for r, h in enumerate(handles) if r != 0]
        self.token = torch.zeros(1, dtype=torch.float32, device="cuda:%d" % ctx.device)

    def __call__(self, want_depth: bool = False):
        if self.ctx.caller_stream != self._torch.cuda.current_stream().cuda_stream:
            self.ctx.synchronize()  # rendering is not on the stream the token travels on: wait for it on the host
        self.dist.all_reduce(self.token)
        out = None
        if self.rank == 0:
            out = self.ctx.resolve_ipc(self.peer_handles, want_depth=want_depth)
        self.dist.all_reduce(self.token)
        return out



class SlicedResolve:
    """The exchange step of the one-process-per-GPU path without NCCL and without a host round trip
    (rzb_resolve_sliced, k_resolve_sliced): stream-ordered behind the render passes of each rank,

      barrier in (flags stored into every peer's exchange header over NVLink) -> rank r sums slice r of ALL ranks'
      float4 accumulators through peer loads, tone-maps it and stores RGBA8 into rank 0's staging image -> barrier out
      (peers may render into their accumulators again; rank 0 copies image + depth to its pinned host buffers).

    Every GPU pulls (N-1)/N of one frame instead of rank 0 pulling N-1 frames, and nothing but the slowest rank's
    render time is on the critical path. All ranks must call it the same number of times (the flags carry the call
    number). `rgba8_pinned` / `depth_pinned`: page-locked numpy arrays on rank 0 (ignored elsewhere); valid after
    wait()."""

    def __init__(self, ctx, rgba8_pinned=None, depth_pinned=None, group=None):
        import torch.distributed as dist
        self.ctx, self.dist = ctx, dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        mine = (ctx.accum_ipc_handle(), ctx.exchange_ipc_handle())
        both = [None] * self.world
        dist.all_gather_object(both, mine, group=group)
        self.accum_handles = [b[0] for b in both]
        self.exchange_handles = [b[1] for b in both]
        self.rgba8, self.depth = rgba8_pinned, depth_pinned
        self._ms = None
        self._pending = False

    def __call__(self):
        root = self.rank == 0
        self.ctx.resolve_sliced(self.rank, self.world, self.accum_handles, self.exchange_handles,
                                self.rgba8 if root else None, self.depth if root else None)
        self._pending = True

    def wait(self):
        """Blocks until this rank's exchange kernel (and, on rank 0, the copies to the host) have finished."""
        if self._pending:
            self._ms = self.ctx.resolve_sliced_wait()
            self._pending = False
        return (self.rgba8, self.depth) if self.rank == 0 else None

    def last_ms(self):
        """Device time of this rank's last exchange kernel (includes waiting for the slowest rank to arrive)."""
        return self._ms
