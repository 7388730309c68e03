"""rt_render_combined / rt_render_multi on ONE GPU (a 1-rank group): partition + render + resolve + tone map + download inside the
library must reproduce the plain rt_render frame -- the single-rank case of Render()'s partition and MPI_Gather (main.cpp:311-319,
345-347; with CommSize == 1 the gather is a copy). The N-rank cases are tests/test_gpu_multi.py (>= 2 GPUs) and bench.py's
`combine_parity` self-check."""
import numpy as np
import pytest

from par_raytracer_b200 import api

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["tiles", "ranges", "samples"])
def test_one_rank_combined_equals_render(golden_scene, mode):
    gs = golden_scene
    S = api.Scene(gs.scene)
    p = gs.params.copy(); p["min_samples"] = p["max_samples"] = gs.render_spp
    want, cnt_w = S.render(gs.cam, p, gs.W, gs.H)
    comm = api.Comm.create(1, 0, None, 0)
    frame, rgba8, luma, cnt = api.render_combined(S, comm, gs.cam, p, gs.W, gs.H, partition=mode, tile=8, want_frame=True, want_rgba8=True)
    assert int(cnt["ray_count"]) == int(cnt_w["ray_count"]) == int(gs.render_counters["ray_count"])
    if mode == "samples":       # sum / n in the resolve kernel == k_finalize's sum / n
        assert np.array_equal(frame.view(np.uint32), want.view(np.uint32))
    else:
        assert np.array_equal(frame.view(np.uint32), want.view(np.uint32))
    assert np.allclose(frame.reshape(-1, 4), gs.render_rgba, rtol=1e-5, atol=1e-6)
    want8, want_luma = api.tonemap(want)
    assert np.array_equal(rgba8, want8) and luma == want_luma
    st = comm.stats()
    assert st["reduce_bytes"] == 0 and not st["peer_memory"]
    # the local (one process) group of one GPU goes through rt_render_multi
    comms = api.Comm.create_local([0])
    frame2, _, _, cnt2 = api.render_multi([S], comms, gs.cam, p, gs.W, gs.H, partition=mode, tile=8)
    assert np.array_equal(frame2.view(np.uint32), want.view(np.uint32)) and int(cnt2["ray_count"]) == int(cnt_w["ray_count"])
    comms[0].close(); comm.close(); S.close()


def test_combined_adaptive_and_argument_errors(golden_scene):
    gs = golden_scene
    S = api.Scene(gs.scene)
    comm = api.Comm.create(1, 0, None, 0)
    p = gs.params.copy(); p["min_samples"], p["max_samples"] = gs.adaptive_minmax
    frame, _, _, cnt = api.render_combined(S, comm, gs.cam, p, gs.W, gs.H, partition="tiles", tile=16, flags=api.RT_FLAG_ADAPTIVE, want_frame=True)
    want, cnt_w = S.render(gs.cam, p, gs.W, gs.H, flags=api.RT_FLAG_ADAPTIVE)
    assert np.array_equal(frame.view(np.uint32), want.view(np.uint32)) and int(cnt["ray_count"]) == int(cnt_w["ray_count"])
    with pytest.raises(api.RtError, match="adaptive"):
        api.render_combined(S, comm, gs.cam, p, gs.W, gs.H, partition="samples", flags=api.RT_FLAG_ADAPTIVE)
    with pytest.raises(api.RtError):
        api.render_combined(S, comm, gs.cam, p, gs.W, gs.H, partition="tiles", tile=0)
    with pytest.raises(api.RtError):
        api.render_combined(S, comm, gs.cam, p, gs.W, gs.H, root=3)
    comm.close(); S.close()


def test_render_is_ordered_after_the_callers_default_stream(golden_scene):
    """rt_render_device with stream handle 0 (torch's default stream is the legacy stream) must wait for work already enqueued there: a
    long chain of writes that poison the output frame is queued on the default stream right before the call; if the render did not wait
    for it, the poison would land after (or in the middle of) the render's own output."""
    import torch
    gs = golden_scene
    S = api.Scene(gs.scene)
    p = gs.params.copy(); p["min_samples"] = p["max_samples"] = gs.render_spp
    want, _ = S.render(gs.cam, p, gs.W, gs.H)
    frame = torch.zeros((gs.W * gs.H, 4), dtype=torch.float32, device="cuda")
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        for _ in range(20):
            big.zero_()                                   # ~1 ms of queued work on the default (legacy) stream ...
        frame.fill_(float("nan"))                         # ... that ends by poisoning the frame
        S.render_device(gs.cam, p, gs.W, gs.H, frame.data_ptr(), flags=api.RT_OUT_MEAN | api.RT_OUT_FULLFRAME, stream=0)
        got = frame.cpu().numpy()                         # default-stream copy: ordered after the render (the library makes the stream wait)
        assert np.array_equal(got.view(np.uint32), want.reshape(-1, 4).view(np.uint32)), f"repetition {rep}: the render overtook the caller's stream"
    S.close()


def test_pinned_output_buffer_gives_the_same_frame(golden_scene):
    """RT_FLAG_PIN_HOST page-locks the caller's buffer for the download; the frame is the same, a second buffer replaces the first
    registration, and a buffer that cannot be registered (a read-only mapping would be one; here: an unaligned view) still works."""
    gs = golden_scene
    S = api.Scene(gs.scene)
    p = gs.params.copy(); p["min_samples"] = p["max_samples"] = gs.render_spp
    n = gs.W * gs.H
    want, cnt = S.render_task(gs.cam, p, gs.W, gs.H)
    for buf in (np.empty((n, 4), np.float32), np.empty((n, 4), np.float32), np.empty(n * 4 + 1, np.float32)[1:].reshape(n, 4)):
        for _ in range(2):
            buf[:] = -1.0
            got, cnt2 = S.render_task(gs.cam, p, gs.W, gs.H, flags=api.RT_OUT_MEAN | api.RT_FLAG_PIN_HOST, out=buf)
            assert got is buf and np.array_equal(buf.view(np.uint32), want.view(np.uint32)) and int(cnt2["ray_count"]) == int(cnt["ray_count"])
    comm = api.Comm.create(1, 0, None, 0)
    out = np.empty((gs.H, gs.W, 4), np.float32)
    frame, _, _, _ = api.render_combined(S, comm, gs.cam, p, gs.W, gs.H, partition="tiles", tile=8, flags=api.RT_FLAG_PIN_HOST, want_frame=True, out=out)
    assert frame is out and np.array_equal(out.reshape(-1, 4).view(np.uint32), want.view(np.uint32))
    comm.close(); S.close()
