"""The multi-GPU combine on real GPUs (needs >= 2 devices; the driver's single-GPU run skips it, `gpurun --gpus 2` runs it):
one process per GPU (rt_render_combined: peer-memory stores through CUDA IPC, or the NCCL reduce of full frames) and one process for
all GPUs (rt_render_multi), tiles / ranges / sample ranges, rank 0's frame against the single-GPU render."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, out_path, ipc=True):
    if not ipc:
        os.environ["RT_B200_NO_IPC"] = "1"           # force the NCCL reduce of full frames
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
    comm = dist.make_comm(rank, rank, world)
    frame, cnt, _ = dist.render_distributed(S, comm, gs.cam, p, gs.W, gs.H, mode=mode, tile=8)
    frame2, cnt2, _ = dist.render_distributed(S, comm, gs.cam, p, gs.W, gs.H, mode=mode, tile=8)       # a second frame into the same buffers
    if rank == 0:
        assert np.array_equal(frame.view(np.uint32), frame2.view(np.uint32)) and int(cnt["ray_count"]) == int(cnt2["ray_count"])
        assert comm.stats()["peer_memory"] == (ipc and mode != "samples"), comm.stats()
    if rank == 0:
        np.save(out_path, frame.reshape(-1, 4))
        np.save(out_path + ".rays.npy", np.array([int(cnt["ray_count"])]))      # rt_render_combined sums the counters on the root
    dist_t.barrier()
    dist_t.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["tiles", "ranges", "samples"])
@pytest.mark.parametrize("ipc", [True, False])
def test_nccl_combine_equals_single_gpu(tmp_path, mode, ipc):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import GoldenScene
    gs = GoldenScene("spheres")
    out = str(tmp_path / "frame.npy")
    world = 2
    mp.spawn(_worker, args=(world, 29700 + (os.getpid() % 1000) + {"tiles": 0, "ranges": 1, "samples": 2}[mode] + (3 if ipc else 0), mode, out, ipc), nprocs=world, join=True)
    frame = np.load(out)
    rays = int(np.load(out + ".rays.npy")[0])
    want = gs.render_rgba
    assert rays == int(gs.render_counters["ray_count"])
    assert np.allclose(frame, want, rtol=1e-5, atol=1e-6)
    assert np.all(frame[:, 3] == 1.0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["tiles", "ranges", "samples"])
@pytest.mark.parametrize("p2p", [True, False])
def test_render_multi_single_process(mode, p2p, monkeypatch):
    """rt_render_multi (one process, 2 GPUs): the peer-memory gather and the NCCL fallback against the single-GPU render of the
    same frame -- bit-exact for the pixel partitions (the gather moves bits), 2e-6 for the sample split (a different sum tree)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import GoldenScene
    from par_raytracer_b200 import api
    gs = GoldenScene("spheres")
    if not p2p:
        monkeypatch.setenv("RT_B200_NO_P2P", "1")
    scenes = [api.Scene(gs.scene, device=d) for d in range(2)]
    comms = api.Comm.create_local([0, 1])
    p = gs.params.copy(); p["min_samples"] = p["max_samples"] = gs.render_spp
    single, cnt1 = scenes[0].render(gs.cam, p, gs.W, gs.H)
    frame, rgba8, luma, cnt = api.render_multi(scenes, comms, gs.cam, p, gs.W, gs.H, partition=mode, tile=8, want_rgba8=True)
    assert comms[0].stats()["peer_memory"] == p2p
    assert int(cnt["ray_count"]) == int(cnt1["ray_count"]) == int(gs.render_counters["ray_count"])
    if mode == "samples":
        assert np.allclose(frame, single, rtol=2e-6, atol=1e-7) and np.all(frame[..., 3] == 1.0)
    else:
        assert np.array_equal(frame.view(np.uint32), single.view(np.uint32))
    want8, want_luma = api.tonemap(frame)
    assert np.array_equal(rgba8, want8) and abs(luma - want_luma) <= 1e-6 * abs(want_luma)
    for c in comms:
        c.close()
    for s in scenes:
        s.close()
