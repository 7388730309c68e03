"""Multi-GPU partition + combine for the render path (one process per GPU, torch.distributed/NCCL).

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
    rank t % world. Inside a tile pixels are listed row by row, so a warp's rays stay coherent."""
    tiles_x = (width + tile - 1) // tile
    tiles_y = (height + tile - 1) // tile
    ids = []
    for t in range(rank, tiles_x * tiles_y, world):
        ty, tx = divmod(t, tiles_x)
        x0, y0 = tx * tile, ty * tile
        xs = np.arange(x0, min(x0 + tile, width), dtype=np.uint32)
        ys = np.arange(y0, min(y0 + tile, height), dtype=np.uint32)
        ids.append((ys[:, None] * np.uint32(width) + xs[None, :]).reshape(-1))
    return np.concatenate(ids).astype(np.uint32) if ids else np.zeros(0, np.uint32)


def sample_partition(total_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """(sample_begin, sample_count) for `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(total_samples, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def combine_frame(local_frame, mode: str, total_samples: int, dst: int = 0, group=None):
    """The collective that replaces MPI_Gather. `local_frame` is a (H*W, 4) float32 torch tensor (cuda+nccl
    or cpu+gloo) holding this rank's pixels (tiles/ranges: resolved colours, zeros elsewhere; samples: raw
    sums, w = samples rendered). After the call rank `dst` holds the finished frame in `local_frame`."""
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


def render_distributed(scene, cam, params, width: int, height: int, mode: str = "tiles", tile: int = 32,
                       rank: Optional[int] = None, world: Optional[int] = None, device=None, frame=None, group=None):
    """Render(cam, scene, w, h) over all ranks of the process group. Returns (frame tensor, counters, gpu_ms).
    The frame is complete on rank 0 only (like Framebuffer.pixels, main.cpp:338-340)."""
    import torch
    import torch.distributed as dist
    from . import api
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    device = device if device is not None else torch.device("cuda", scene.device)
    if frame is None:
        frame = torch.zeros((width * height, 4), dtype=torch.float32, device=device)
    else:
        frame.zero_()
    spp = int(np.asarray(params)["min_samples"])
    stream = torch.cuda.current_stream(device).cuda_stream
    if mode == "tiles":
        ids = tile_partition(width, height, rank, world, tile)
        cnt = scene.render_device(cam, params, width, height, frame.data_ptr(), pixel_ids=ids, sample_count=spp,
                                  flags=api.RT_OUT_MEAN | api.RT_OUT_FULLFRAME, stream=stream)
    elif mode == "ranges":
        start, count = range_partition(width, height, rank, world)
        cnt = scene.render_device(cam, params, width, height, frame.data_ptr(), pixel_begin=start, pixel_count=count,
                                  sample_count=spp, flags=api.RT_OUT_MEAN | api.RT_OUT_FULLFRAME, stream=stream)
    elif mode == "samples":
        s0, ns = sample_partition(spp, rank, world)
        cnt = scene.render_device(cam, params, width, height, frame.data_ptr(), pixel_begin=0, pixel_count=width * height,
                                  sample_begin=s0, sample_count=ns, flags=api.RT_OUT_SUM | api.RT_OUT_FULLFRAME, stream=stream)
    else:
        raise ValueError(f"unknown mode {mode!r}")
    gpu_ms = float(scene.stats()["gpu_ms"])
    combine_frame(frame, mode, spp, dst=0, group=group)
    return frame, cnt, gpu_ms
