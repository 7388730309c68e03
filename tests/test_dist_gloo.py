"""world_size-2 (and 3) gloo runs of the multi-GPU combine on CPU tensors: every rank renders its share with
the oracle (stand-in for its GPU), calls the SAME partition + combine code the NCCL path uses, and rank 0's
frame must equal the single-process render -- bit-exact for tiles / ranges, tolerance for the sample split."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist_t
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist_t.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import GoldenScene
    from oracle import oracle
    from par_raytracer_b200 import dist
    gs = GoldenScene("spheres")
    O = oracle.OracleScene(gs.scene)
    W, H, spp = gs.W, gs.H, gs.render_spp
    p = gs.params.copy(); p["min_samples"] = p["max_samples"] = spp
    frame = torch.zeros((W * H, 4), dtype=torch.float32)
    if mode == "tiles":
        ids = dist.tile_partition(W, H, rank, world, tile=8)
        img, _, _, _ = O.render(gs.cam, p, W, H, pixel_ids=ids)
        frame[torch.from_numpy(ids.astype(np.int64))] = torch.from_numpy(img)
    elif mode == "ranges":
        start, count = dist.range_partition(W, H, rank, world)
        img, _, _, _ = O.render(gs.cam, p, W, H, pixel_begin=start, pixel_count=count)
        frame[start:start + count] = torch.from_numpy(img)
    else:
        s0, ns = dist.sample_partition(spp, rank, world)
        q = p.copy(); q["min_samples"] = q["max_samples"] = ns
        img, _, _, _ = O.render(gs.cam, q, W, H, sample_begin=s0, sum_only=True)
        frame[:] = torch.from_numpy(img)
        frame[:, 3] = ns
    dist.combine_frame(frame, mode, spp, dst=0)
    if rank == 0:
        np.save(out_path, frame.numpy())
    dist_t.barrier()
    dist_t.destroy_process_group()


@pytest.mark.parametrize("mode,world", [("tiles", 2), ("ranges", 2), ("samples", 2), ("tiles", 3)])
def test_combine_equals_single_process(tmp_path, mode, world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import GoldenScene
    gs = GoldenScene("spheres")
    out = str(tmp_path / "frame.npy")
    port = 29500 + (os.getpid() % 2000) + {"tiles": 0, "ranges": 1, "samples": 2}[mode] + 7 * world
    mp.spawn(_worker, args=(world, port, mode, out), nprocs=world, join=True)
    frame = np.load(out)
    want = gs.render_rgba                      # the reference's own seeded render of this frame
    if mode == "samples":
        assert np.allclose(frame, want, rtol=2e-6, atol=1e-7)     # same samples, different summation tree
        assert np.all(frame[:, 3] == 1.0)
    else:
        assert np.array_equal(frame.view(np.uint32), want.view(np.uint32))   # x + 0 == x: reduce == MPI_Gather
