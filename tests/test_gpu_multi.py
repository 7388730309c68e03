"""NCCL combine on real GPUs (needs >= 2 devices; the driver's single-GPU run skips it, `gpurun --gpus 2` runs it):
one process per GPU, tiles / ranges / sample ranges, rank 0's frame against the single-GPU render."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist_t
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist_t.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from conftest import GoldenScene
    from par_raytracer_b200 import api, dist
    gs = GoldenScene("spheres")
    S = api.Scene(gs.scene, device=rank)
    p = gs.params.copy(); p["min_samples"] = p["max_samples"] = gs.render_spp
    frame, cnt, _ = dist.render_distributed(S, gs.cam, p, gs.W, gs.H, mode=mode, tile=8, rank=rank, world=world)
    rays = torch.tensor([int(cnt["ray_count"])], dtype=torch.int64, device=f"cuda:{rank}")
    dist_t.all_reduce(rays)
    if rank == 0:
        np.save(out_path, frame.cpu().numpy())
        np.save(out_path + ".rays.npy", rays.cpu().numpy())
    dist_t.barrier()
    dist_t.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["tiles", "ranges", "samples"])
def test_nccl_combine_equals_single_gpu(tmp_path, mode):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import GoldenScene
    gs = GoldenScene("spheres")
    out = str(tmp_path / "frame.npy")
    world = 2
    mp.spawn(_worker, args=(world, 29700 + (os.getpid() % 1000) + {"tiles": 0, "ranges": 1, "samples": 2}[mode], mode, out), nprocs=world, join=True)
    frame = np.load(out)
    rays = int(np.load(out + ".rays.npy")[0])
    want = gs.render_rgba
    assert rays == int(gs.render_counters["ray_count"])
    assert np.allclose(frame, want, rtol=1e-5, atol=1e-6)
    assert np.all(frame[:, 3] == 1.0)
