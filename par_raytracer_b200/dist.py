"""Multi-GPU partition + combine for the render path, one process per GPU. The product path is the C ABI
(rt_comm_* / rt_render_combined: NCCL C API on the render stream + resolve kernel); torch.distributed only carries the
128-byte NCCL unique id between the processes. `combine_frame` is the host-side model used by the gloo CPU tests.

The reference shards contiguous pixel-index ranges over MPI ranks with the scene replicated and gathers
the ranges on rank 0 (main.cpp:311-319, 345-347). Here:

  * mode "tiles"   -- interleaved square tiles, round-robin over ranks (contiguous ranges load-balance
                      badly: NOTES.txt:25, out.txt shows 2x imbalance). Every rank writes its pixels into a
                      zero-initialised full frame; reduce(SUM) to rank 0 is then bit-identical to the
                      reference's MPI_Gather (x + 0 == x).
  * mode "ranges"  -- the reference's own partition: rank r renders [r*cpp, (r+1)*cpp), cpp = ceil(W*H / ranks).
  * mode "samples" -- every rank renders all pixels for a sample sub-range [s0, s1) as raw sums
                      (RT_OUT_SUM); reduce(SUM), then divide by the total sample count and set w = 1.

There is no exchange inside the path itself: the single collective is the frame combine.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def range_partition(width: int, height: int, rank: int, world: int) -> Tuple[int, int]:
    """main.cpp:311-317: (start_idx, count) of the reference's contiguous split, clipped to the frame."""
    total = width * height
    cpp = (total + world - 1) // world
    start = min(total, cpp * rank)
    end = min(total, cpp * (rank + 1))
    return start, end - start


def tile_partition(width: int, height: int, rank: int, world: int, tile: int = 32) -> np.ndarray:
    """Linear pixel ids (row-major, uint32) of the tiles owned by `rank`: tile t = ty * tiles_x + tx goes to
    rank t % world; inside a tile pixels are listed row by row, so a warp's rays stay coherent. Computed by the
    library's own host function (rt_partition_tiles), the one rt_render_combined uses."""
    from . import api
    return api.partition_tiles(width, height, rank, world, tile)


def sample_partition(total_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """(sample_begin, sample_count) for `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(total_samples, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def combine_frame(local_frame, mode: str, total_samples: int, dst: int = 0, group=None):
    """Host-side MODEL of the combine for CPU (gloo) runs of the partition logic: reduce(SUM) to `dst`, then the resolve the
    partition implies. On GPUs the product path is rt_render_combined (NCCL C API + resolve kernel inside the library)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(local_frame, dst=dst, op=dist.ReduceOp.SUM, group=group)
        is_dst = dist.get_rank(group) == dst
    else:
        is_dst = True
    if mode == "samples" and is_dst:
        local_frame[:, :3] /= float(total_samples)       # main.cpp:262
        local_frame[:, 3] = 1.0                           # main.cpp:263
    return local_frame


def make_comm(device: int, rank: Optional[int] = None, world: Optional[int] = None, group=None):
    """rt_comm for this rank of the torch.distributed job: rank 0 draws the NCCL unique id (rt_comm_unique_id), the process
    group's own broadcast hands it out (an MPI host would use one MPI_Bcast), every rank calls rt_comm_create."""
    import torch
    import torch.distributed as dist
    from . import api
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    uid = None
    if world > 1:
        box = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        uid = box[0]
    return api.Comm.create(world, rank, uid, device)


def render_distributed(scene, comm, cam, params, width: int, height: int, mode: str = "tiles", tile: int = 32, flags: int = 0,
                       want_frame: bool = True):
    """Render(cam, scene, w, h) over all ranks of `comm`: one rt_render_combined call per rank. Returns (frame (H, W, 4) float32
    on rank 0 / None elsewhere, counters, gpu_ms of this rank's render). Like Framebuffer.pixels the frame exists on rank 0 only
    (main.cpp:338-340)."""
    from . import api
    frame, _, _, cnt = api.render_combined(scene, comm, cam, params, width, height, partition=mode, tile=tile, flags=flags, root=0,
                                           want_frame=want_frame)
    return frame, cnt, float(scene.stats()["gpu_ms"])
