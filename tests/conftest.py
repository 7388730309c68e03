import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class GoldenScene:
    """A tests/golden/scene_*.npz file: scene exported by the reference + the reference's outputs on it."""

    def __init__(self, name):
        from par_raytracer_b200.types import CAMERA, COUNTERS, HIT, PARAMS, RAY, SceneData
        z = np.load(os.path.join(GOLDEN, f"scene_{name}.npz"))
        self.z = z
        self.scene = SceneData.from_npz_dict(z, name=name)
        self.cam = np.frombuffer(z["cam"].tobytes(), CAMERA)[0]
        self.params = np.frombuffer(z["params"].tobytes(), PARAMS)[0].copy()
        self.W, self.H = (int(v) for v in z["wh"])
        self.primary_rays = np.frombuffer(z["primary_rays"].tobytes(), RAY)
        self.primary_hits = np.frombuffer(z["primary_hits"].tobytes(), HIT)
        self.random_rays = np.frombuffer(z["random_rays"].tobytes(), RAY)
        self.random_hits = np.frombuffer(z["random_hits"].tobytes(), HIT)
        self.random_counters = np.frombuffer(z["random_counters"].tobytes(), COUNTERS)[0]
        self.color_rays = np.frombuffer(z["color_rays"].tobytes(), RAY)
        self.color_seeds = z["color_seeds"]
        self.color_rgba = z["color_rgba"]
        self.color_counters = np.frombuffer(z["color_counters"].tobytes(), COUNTERS)[0]
        self.render_spp = int(z["render_spp"][0])
        self.render_rgba = z["render_rgba"]
        self.render_counters = np.frombuffer(z["render_counters"].tobytes(), COUNTERS)[0]
        self.render_sum_from2 = z["render_sum_from2"]
        self.adaptive_minmax = tuple(int(v) for v in z["adaptive_minmax"])
        self.adaptive_rgba = z["adaptive_rgba"]
        self.adaptive_nsamples = z["adaptive_nsamples"]
        self.adaptive_counters = np.frombuffer(z["adaptive_counters"].tobytes(), COUNTERS)[0]
        self.tonemap_rgba8 = z["tonemap_rgba8"]
        self.tonemap_luma = float(z["tonemap_luma"][0])


@pytest.fixture(scope="session", params=["spheres", "heightfield"])
def golden_scene(request):
    return GoldenScene(request.param)


@pytest.fixture(scope="session")
def golden_functions():
    return np.load(os.path.join(GOLDEN, "functions.npz"))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_hits_equal(got, want, what=""):
    """Bit-exact comparison of rt_hit arrays in every field the reference defines (on a miss only `hit`,
    `object` and t == FLT_MAX are defined: raytracer.cpp:166, 230)."""
    assert np.array_equal(got["hit"], want["hit"]), f"{what}: hit mask differs at {np.flatnonzero(got['hit'] != want['hit'])[:8]}"
    assert np.array_equal(got["object"], want["object"]), f"{what}: object (leaf sphere index) differs"
    m = want["hit"] == 1
    assert np.array_equal(got["vertex0"][m], want["vertex0"][m]), f"{what}: vertex0 differs"
    assert np.array_equal(bits(got["t"]), bits(want["t"])), f"{what}: t differs"
    for f in ("bw", "position", "normal"):
        assert np.array_equal(bits(got[f][m]), bits(want[f][m])), f"{what}: {f} differs"
