#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (Mrays/s, primary + secondary; ms/frame) on BASELINE config 3:
procedural OBJ-style mesh, 1,048,352 textured triangles in 529 mesh groups (diffuse / ambient / bump / alpha maps), 1920x1080, 128 spp
fixed, reference defaults (bounce_depth 2, 1 diffuse + 1 specular sample, 1 directional light). It is the largest BASELINE
configuration whose reference arm is still feasible on host cores through the reference's own OBJ parser and BuildHierarchy, and
its scene (16 MB of nodes + 48 MB of triangle records + shading records) no longer lives in L1 like config 2's 4 MB scene does.

A "step" is one Render() of that frame. A "ray" is one TraceRay call (raytracer.cpp:161): primary, shadow, diffuse / specular bounce
and alpha continuation rays. The timed region is the reference's "Render, sync" block (main.cpp:326-333) plus, for N > 1, its "Reduce"
block (main.cpp:343-348).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload configK] [--partition tiles|samples|ranges]
                    [--also config2,config4 | none]

N = 1: `value` = rt_render_device (frame stays in HBM); `e2e` = rt_render, the host-buffer call of the C ABI (host camera / params in,
pageable host framebuffer out) with the same flags. `also` carries the same measurement of BASELINE config 2 (the L1/L2-resident scene)
and config 4 (10 M triangles, 3840x2160, 256 spp) taken in the same process.

N > 1 (torchrun, one rank per GPU): STRONG scaling of the same frame -- interleaved 32x32 tiles per rank (the reference's pixel split,
main.cpp:311-319, made interleaved: NOTES.txt:25), one rt_render_combined call per rank and step: render + the gather of the float4 frames on
rank 0 (replaces MPI_Gather, main.cpp:345-347) inside the library -- peer-memory stores into rank 0's frame over NVLink where CUDA IPC maps it,
else ncclReduce(SUM) on the render stream. `combine_parity`: untimed check that
rank 0's combined frame is bit-identical to its own single-GPU render of the whole frame.

--impl reference: times the reference's own CPU implementation (oracle/_ref = the unmodified reference compiled in the authoring
container; else the C oracle port) on all host cores, on a bounded pixel subset of the same frame (the metric is a rate).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (primary+secondary)"
DEFAULT_WORKLOAD = "config3"
TILE = int(os.environ.get("RT_BENCH_TILE", "32"))      # interleaved tile edge of the N > 1 partition (measured 8 / 16 / 32 / 64 at N = 8: profiles/README.md)

# key -> (width, height, spp, description)
WORKLOADS = {
    "config1": (720, 480, 10, "config1: reference defaults (main.cpp:419-434, 308-309) on the Sponza stand-in: 720x480, adaptive 10..50 spp"),
    "config2": (1920, 1080, 64, "config2: 16 tessellated spheres + plane, 63490 triangles / 17 groups, 1920x1080, 64 spp, bounce_depth 2"),
    "config3": (1920, 1080, 128, "config3: procedural height field, 1048352 textured triangles / 529 groups (diffuse, ambient, bump, alpha maps), 1920x1080, 128 spp, bounce_depth 2"),
    "config4": (3840, 2160, 256, "config4: procedural height field, 9999392 triangles / 4900 groups, 3840x2160, 256 spp, bounce_depth 2"),
    "config5": (1920, 1080, 4096, "config5: config-3 scene (1048352 textured triangles), 1920x1080, 4096 spp split by sample ranges over the GPUs"),
}

ROOFLINE_NOTE = {
    "config1": "adaptive sampling: many small waves; not a roofline workload",
    "config2": "the 4 MB scene (1.1 MB nodes + 3 MB triangle records) is L1/L2-resident: measured DRAM traffic per ray is far BELOW the algorithmic 776 B "
               "and frac can exceed 1; on this scene the kernel is ALU-pipe bound (ncu, profiles/). Reported for comparison, not as a roofline claim",
    "config3": "achieved counts SURVEY 8(d)'s algorithmic bytes (904 B per ray); the 64 MB of nodes + triangle records exceed L1 and compete with the "
               "path-state streams for L2: the kernel waits on scattered node / triangle loads (ncu, profiles/)",
    "config5": "achieved counts SURVEY 8(d)'s algorithmic bytes (904 B per ray)",
    "config4": "achieved counts SURVEY 8(d)'s algorithmic bytes (1032 B per ray); 161+ MB of nodes + 480 MB of triangle records exceed the 126 MB L2: "
               "latency of dependent scattered loads from L2 / HBM bounds the kernel (ncu, profiles/)",
}


def b_ray(n_tris: int) -> int:
    """algorithmic bytes per ray = 64 (ray w+r) + 32 (hit w+r) + 32 * ceil(log2(N/4)) (nodes)
    + 4 * 36 (leaf triangles) + 88 (amortised shading traffic)."""
    return 64 + 32 + 32 * math.ceil(math.log2(max(2.0, n_tris / 4.0))) + 144 + 88


def measured_traffic(workload: str, kernel: str):
    """(DRAM bytes per launch, source file) of `kernel` from the committed ncu pass of this workload, or (None, None). The number is NOT
    measured in this run (ncu replays kernels; a bench value is never taken under it): the source file is named next to it."""
    for fn in (f"r2_traffic_{workload}.json", "r1_traffic_per_kernel.json" if workload == "config2" else None):
        if not fn:
            continue
        p = os.path.join(ROOT, "profiles", fn)
        try:
            k = json.load(open(p))["kernels"]
            for name, v in k.items():
                if name.startswith(kernel):
                    return float(v["dram_bytes_per_launch"]), "profiles/" + fn
        except Exception:
            pass
    return None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe), read through NVML from a
    background thread (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; a polling
    nvidia-smi process was measured to slow the many-sync wave loop by ~40 ms per step, NVML reads do not)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, interval_s: float = 0.1):
        self.index, self.interval = index, interval_s
        self.sm, self.mask, self.power = [], 0, []
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_sm = None
        self.mode = os.environ.get("RT_BENCH_CLOCKS", "nvml")

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")       # torch's device order follows CUDA_VISIBLE_DEVICES; NVML's does not
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            self.stop_flag.wait(self.interval)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.sm.append(float(out[0])); self.max_sm = float(out[1])
                for k, b in enumerate(bits):
                    if out[2 + k].strip().lower().startswith("active"):
                        self.mask |= b
            except Exception:
                pass
            self.stop_flag.wait(self.interval)

    def start(self):
        if self.mode == "off":
            return self
        self.thread = threading.Thread(target=self._run_nvml if self.mode == "nvml" else self._run_smi, daemon=True)
        self.thread.start()
        return self

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=5)
        reasons = sorted(n for b, n in self.REASONS.items() if self.mask & b)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None, "source": self.mode}


# -------------------------------------------------------------------------------------------------
# inputs
# -------------------------------------------------------------------------------------------------
def make_scene(key: str):
    """Host-side scene of a BASELINE configuration (synthetic, seeded generators in par_raytracer_b200/scenes.py). The height-field
    scenes defer their group hierarchy: the product arm takes the reference's own BuildHierarchy from rt_build_group_hierarchy (GPU,
    bit-identical), the reference arm runs the reference's BuildHierarchy on the OBJ -- the same hierarchy on both sides."""
    from par_raytracer_b200 import scenes
    if key in ("config3", "config5"):
        return scenes.heightfield_scene(724, 724, block=32, size=400.0, amp=20.0, textured=True, tex_size=512, hierarchy="defer")
    if key == "config4":
        return scenes.heightfield_scene(2236, 2236, block=32, size=400.0, amp=20.0, textured=False, hierarchy="defer")
    if key == "config1":
        return scenes.sponza_standin_scene()
    return scenes.spheres_plane_scene()


def make_camera_params(key: str, sd):
    from par_raytracer_b200 import types
    W, H, spp, _ = WORKLOADS[key]
    h = sd.camera_hint
    cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
    params = types.default_params(spp=spp, max_spp=50 if key == "config1" else None)
    return cam, params


# -------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (or the oracle port) on the host cores
# -------------------------------------------------------------------------------------------------
class CpuArm:
    """The seeded CPU render of a pixel subset of the workload's frame (every k-th pixel in x and y), on all host cores. The sample is
    sized by a pilot so that one step is about `budget_s` seconds: first in spp (up to the workload's), then in pixels."""

    def __init__(self, key: str, sd, cam, params):
        from oracle import oracle, ref_harness
        self.key, self.cam, self.params = key, cam, params
        self.W, self.H, self.spp, _ = WORKLOADS[key]
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.seed = int(params["base_seed"])
        self.adaptive = int(params["min_samples"]) < int(params["max_samples"])
        if ref_harness.available():
            self.kind = "reference"
            from par_raytracer_b200 import scenes
            d = tempfile.mkdtemp(prefix="bench_ref_scene_")
            scenes.write_obj(sd, d)                       # the reference parses the OBJ and runs its own BuildHierarchy
            self.R = ref_harness.get()
            self.R.load_scene(d)
            self.R.set_params(params)
            self.R.set_lights(sd.lights)
        else:
            self.kind = "port"
            if not sd.spheres.size:
                from par_raytracer_b200 import scenes
                sd.spheres, sd.sphere_group = scenes.build_group_hierarchy(sd.positions, sd.group_first, sd.idx_positions)
            self.O = oracle.OracleScene(sd)

    def subset(self, stride):
        xs = np.arange(stride // 2, self.W, stride, dtype=np.uint32)
        ys = np.arange(stride // 2, self.H, stride, dtype=np.uint32)
        return (ys[:, None] * np.uint32(self.W) + xs[None, :]).reshape(-1).astype(np.uint32)

    def run(self, ids, spp):
        lo, hi = (spp, spp) if not self.adaptive else (int(self.params["min_samples"]), int(self.params["max_samples"]))
        if self.kind == "reference":
            _, _, cnt, sec = self.R.render_seeded(self.cam, self.W, self.H, ids, 0, len(ids), 0, lo, hi, self.seed, threads=self.cores)
        else:
            p = self.params.copy(); p["min_samples"], p["max_samples"] = lo, hi
            _, _, cnt, sec = self.O.render(self.cam, p, self.W, self.H, pixel_ids=ids, threads=self.cores)
        return int(cnt["ray_count"]), sec

    def measure(self, budget_s: float, steps: int = 1, warmup: int = 0) -> dict:
        stride = 48
        ids = self.subset(stride)
        pilot_spp = 1 if not self.adaptive else self.spp
        _, sec1 = self.run(ids, pilot_spp)
        spp = self.spp if self.adaptive else int(max(1, min(self.spp, round(budget_s / max(sec1, 1e-6)))))
        if spp == self.spp:
            per_spp = sec1 / pilot_spp
            for cand in (32, 24, 16, 12, 8, 6, 4):
                if per_spp * self.spp * (48.0 / cand) ** 2 <= budget_s * 1.25:
                    stride = cand
            ids = self.subset(stride)
        for _ in range(warmup):
            self.run(ids, spp)
        tot_r, tot_s = 0, 0.0
        for _ in range(max(1, steps)):
            r, s = self.run(ids, spp)
            tot_r += r; tot_s += s
        return {"value": tot_r / tot_s / 1e6, "unit": "Mrays/s", "cores": self.cores, "kind": self.kind,
                "sample": f"{len(ids)} pixels (every {stride}th in x and y of {self.W}x{self.H}) x {spp if not self.adaptive else 'adaptive 10..50'} spp, "
                          f"{tot_r // max(1, steps)} rays/step, per-(pixel,sample) seeding, {self.cores} threads over contiguous pixel chunks",
                "seconds_per_step": tot_s / max(1, steps), "rays": tot_r}


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    key = args.workload
    sd = make_scene(key)
    cam, params = make_camera_params(key, sd)
    arm = CpuArm(key, sd, cam, params)
    # the whole `--steps K --warmup W` run should end within a few minutes: ~200 s of timed + warm-up CPU work
    budget = min(20.0, max(1.0, 200.0 / max(1, args.steps + args.warmup)))
    r = arm.measure(budget_s=budget, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[key][3], "note": "CPU reference path: the unmodified reference on the host cores renders a bounded pixel subset of the same frame; the metric is a rate"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------------
# product arm
# -------------------------------------------------------------------------------------------------
class Ctx:
    pass


def measure_workload(ctx, key: str, steps: int, warmup: int, partition: str, with_clocks: bool, parity_check: bool):
    """value / e2e / roofline of one workload on ctx's ranks. Returns the dict of line fields (rank 0) or None."""
    import torch
    import torch.distributed as dist_t
    from par_raytracer_b200 import api, scenes
    rank, world, dev, local_rank = ctx.rank, ctx.world, ctx.dev, ctx.local_rank
    W, H, spp, desc = WORKLOADS[key]
    t0 = time.time()
    sd = make_scene(key)
    gen_s = time.time() - t0
    if not sd.spheres.size:
        scenes.use_reference_hierarchy(sd, device=local_rank)        # the reference's BuildHierarchy, on the GPU (bit-identical)
    cam, params = make_camera_params(key, sd)
    adaptive = int(params["min_samples"]) < int(params["max_samples"])
    t0 = time.time()
    S = api.Scene(sd, device=local_rank)
    scene_create_s = time.time() - t0
    info = S.hierarchy_info()
    n_tris = info["triangles"]
    n_px = W * H
    base_flags = api.RT_FLAG_TIME_KERNELS | (api.RT_FLAG_ADAPTIVE if adaptive else 0)
    frame = torch.zeros((n_px, 4), dtype=torch.float32, device=dev) if world == 1 else None
    host_frame = np.empty((n_px, 4), np.float32)                     # plain host memory, what a C host hands to rt_render
    cam_h = np.asarray(cam).reshape(1).copy(); par_h = np.asarray(params).reshape(1).copy()

    def sync_all():
        if world > 1:
            dist_t.barrier()
        torch.cuda.synchronize(dev)

    def step(e2e: bool):
        """One Render() of the frame."""
        ctx.flush.zero_()                     # > L2; the render is ordered after it (legacy-stream wait inside the library)
        if world == 1:
            if e2e:
                _, cnt = S.render_task(cam_h[0], par_h[0], W, H, 0, n_px, flags=api.RT_OUT_MEAN | base_flags | api.RT_FLAG_PIN_HOST, out=host_frame)
                float(host_frame[0, 0])       # the caller reads the result
            else:
                cnt = S.render_device(cam, params, W, H, frame.data_ptr(), flags=api.RT_OUT_MEAN | api.RT_OUT_FULLFRAME | base_flags, stream=0)
            cst = None
        else:
            out, _, _, cnt = api.render_combined(S, ctx.comm, cam_h[0], par_h[0], W, H, partition=partition, tile=TILE, flags=base_flags | (api.RT_FLAG_PIN_HOST if e2e else 0), root=0,
                                                 want_frame=e2e, out=host_frame if (e2e and rank == 0) else None)
            if e2e and rank == 0:
                float(host_frame[0, 0])
            cst = ctx.comm.stats()
        return int(cnt["ray_count"]), S.stats(), cst

    def timed_loop(e2e: bool):
        for _ in range(warmup):
            step(e2e)
        sync_all()
        sampler = ClockSampler(local_rank).start() if (with_clocks and rank == 0 and not e2e) else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = dict(rays=0, launches=0, trace_ms=0.0, logic_ms=0.0, closest=0, shadow=0, waves=0, gpu_ms=0.0, combine_ms=0.0, deliver_ms=0.0)
        sync_all()
        ev0.record()
        for _ in range(steps):
            r, st, cst = step(e2e)
            acc["rays"] += r if (world == 1 or rank == 0) else 0       # N > 1: rank 0's counters are already the sum over ranks
            acc["launches"] += int(st["kernel_launches"]) + (1 if world > 1 else 0)
            acc["trace_ms"] += float(st["trace_ms"]); acc["logic_ms"] += float(st["logic_ms"]); acc["gpu_ms"] += float(st["gpu_ms"])
            acc["closest"] += int(st["closest_rays"]); acc["shadow"] += int(st["shadow_rays"]); acc["waves"] += int(st["waves"])
            if cst:
                acc["combine_ms"] += cst["combine_ms"]; acc["deliver_ms"] += cst["deliver_ms"]
        ev1.record()
        sync_all()
        acc["ms"] = ev0.elapsed_time(ev1)
        acc["clocks"] = sampler.stop() if sampler else None
        return acc

    dev_run = timed_loop(False)
    e2e_run = timed_loop(True)

    # ---- untimed: N > 1 parity of the combined frame against a single-GPU render of the same frame on rank 0 ----
    combine_parity = None
    combine_max_rel_diff = 0.0
    if world > 1 and parity_check:
        out, _, _, cnt = api.render_combined(S, ctx.comm, cam_h[0], par_h[0], W, H, partition=partition, tile=TILE, flags=base_flags & ~api.RT_FLAG_TIME_KERNELS,
                                             root=0, want_frame=True)
        if rank == 0:
            single, cnt1 = S.render(cam_h[0], par_h[0], W, H, flags=api.RT_OUT_MEAN | (api.RT_FLAG_ADAPTIVE if adaptive else 0))
            if partition == "samples":
                # a different summation tree (N partial sums of spp / N samples instead of one running sum of spp): float32 rounding differs by
                # ~sqrt(spp) * 2^-24 relative -- observed 2.7e-5 at 4096 spp, bound 1e-6 * sqrt(spp); the ray counts must be equal
                rel = np.abs(out - single) / np.maximum(np.abs(single), 1e-6)
                combine_max_rel_diff = float(rel.max())
                combine_parity = combine_max_rel_diff <= max(2e-6, 1e-6 * math.sqrt(float(params["min_samples"]))) and int(cnt["ray_count"]) == int(cnt1["ray_count"])
            else:
                combine_parity = bool(np.array_equal(out.view(np.uint32), single.view(np.uint32))) and int(cnt["ray_count"]) == int(cnt1["ray_count"])
        sync_all()

    if world > 1:
        t = torch.tensor([dev_run["ms"], e2e_run["ms"], dev_run["trace_ms"], dev_run["logic_ms"], dev_run["gpu_ms"], dev_run["combine_ms"]], dtype=torch.float64, device=dev)
        dist_t.all_reduce(t, op=dist_t.ReduceOp.MAX)
        dev_run["ms"], e2e_run["ms"] = float(t[0]), float(t[1])
        max_trace_ms, max_logic_ms, max_gpu_ms, max_combine_ms = float(t[2]), float(t[3]), float(t[4]), float(t[5])
        c = torch.tensor([dev_run["rays"], e2e_run["rays"], dev_run["launches"], dev_run["closest"] + dev_run["shadow"], dev_run["waves"]], dtype=torch.float64, device=dev)
        dist_t.all_reduce(c, op=dist_t.ReduceOp.SUM)
        rays, e_rays, launches, traced_all, waves_all = int(c[0]), int(c[1]), int(c[2]), int(c[3]), int(c[4])
    else:
        rays, e_rays, launches = dev_run["rays"], e2e_run["rays"], dev_run["launches"]
        traced_all, waves_all = dev_run["closest"] + dev_run["shadow"], dev_run["waves"]
        max_trace_ms, max_logic_ms, max_gpu_ms, max_combine_ms = dev_run["trace_ms"], dev_run["logic_ms"], dev_run["gpu_ms"], 0.0
    S.close()
    if rank != 0:
        return None

    peak, peak_src = measured_peak_gbs()
    if world == 1:
        combine_desc = "none"
    elif ctx.comm.stats()["peer_memory"]:
        combine_desc = ("rt_render_combined, inside the timed region: every rank's resolve kernel stores its tiles straight into rank 0's %d MB frame over NVLink "
                        "(CUDA IPC peer memory: no zero-fill, no reduce of frames); one ncclReduce of the 24-byte counters is the completion signal" % (n_px * 16 >> 20))
    else:
        combine_desc = "rt_render_combined, inside the timed region: ncclReduce(SUM) of the %d MB float4 frames to rank 0 on the render stream" % (n_px * 16 >> 20)
    bpr = b_ray(n_tris)
    ms, e_ms = dev_run["ms"], e2e_run["ms"]
    value = rays / (ms * 1e-3) / 1e6
    # dominant kernel: k_trace_wave traces the closest-hit rays and the shadow rays of a wave in one launch. Rank 0's launches, rank 0's rays.
    traced0 = dev_run["closest"] + dev_run["shadow"]
    achieved = traced0 * bpr / (dev_run["trace_ms"] * 1e-3) / 1e9 if dev_run["trace_ms"] > 0 else 0.0
    traffic, traffic_src = measured_traffic(key, "k_trace_wave") if world == 1 else (None, None)
    mode = "none" if world == 1 else partition
    out = {
        "value": value, "ms_per_step": ms / steps,
        "config": {"workload": desc, "width": W, "height": H, "spp_total": int(params["max_samples"]) if adaptive else spp,
                   "partition": mode, "tile": TILE if mode == "tiles" else None,
                   "combine": combine_desc,
                   "l2": "256 MB memset flushes L2 before every step; per-step path-state streams (tens of GB) exceed the 126 MB L2",
                   "hierarchy": "reference BuildHierarchy over the mesh groups (rt_build_group_hierarchy, bit-identical) for the tie-break order; traversal on the GPU-built cluster hierarchy",
                   "triangles": n_tris, "hierarchy_nodes": info["nodes"], "scene_generate_s": gen_s, "scene_create_s": scene_create_s,
                   "rays_per_step": rays // steps, "waves_per_step": dev_run["waves"] // steps,
                   "kernel_ms_per_step": {"k_trace_wave": dev_run["trace_ms"] / steps, "k_logic": dev_run["logic_ms"] / steps},
                   "steps": steps, "warmup": warmup},
        "e2e": {"value": e_rays / (e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e_ms / steps,
                "h2d_bytes_per_step": int(cam_h.nbytes + par_h.nbytes), "d2h_bytes_per_step": int(n_px * 16),
                "note": ("rt_render: host camera / params in, the caller's host framebuffer out (a plain malloc'ed buffer the library page-locks on first sight, RT_FLAG_PIN_HOST), same flags as `value`" if world == 1 else
                         "rt_render_combined with the finished frame downloaded to rank 0's host buffer every step (page-locked by the library on first sight, RT_FLAG_PIN_HOST)") +
                        "; the scene stays resident like the reference's loaded Scene"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "k_trace_wave", "bytes_per_ray": bpr, "rays_timed": traced0, "kernel_ms": dev_run["trace_ms"],
                     "launches_timed": dev_run["waves"], "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": traced0 * bpr / max(1, dev_run["waves"]),
                     "note": ROOFLINE_NOTE[key],
                     "whole_step_frac": value * 1e6 / world * bpr / (peak * 1e9)},
        "clocks": dev_run["clocks"],
    }
    if world > 1:
        out["combine_parity"] = combine_parity
        out["combine_max_rel_diff"] = combine_max_rel_diff
        out["per_step_ms"] = {"render_gpu_max_over_ranks": max_gpu_ms / steps, "k_trace_wave_max": max_trace_ms / steps, "k_logic_max": max_logic_ms / steps,
                              "reduce_and_resolve_max": max_combine_ms / steps,
                              "note": "render = CUDA-event time of the rank's wave loop; the wave tails (launch + drain of ~25-30 waves x 2 kernels) do not shrink with the tile count"}
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist_t
    from par_raytracer_b200 import api, dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.rank, ctx.world, ctx.local_rank = rank, world, local_rank
    ctx.dev = torch.device("cuda", local_rank)
    if world > 1:
        dist_t.init_process_group("nccl", device_id=ctx.dev)
    ctx.comm = dist.make_comm(local_rank, rank, world) if world > 1 else None       # NCCL unique id from rank 0 over the process group
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)            # > 126 MB L2
    key = args.workload
    partition = args.partition or ("samples" if key == "config5" else "tiles")

    head = measure_workload(ctx, key, args.steps, args.warmup, partition, with_clocks=True, parity_check=True)
    also = {}
    also_keys = [k for k in (args.also.split(",") if args.also != "none" else []) if k and k != key]
    for k in also_keys:
        if k not in WORKLOADS:
            continue
        st, wu = (min(args.steps, 3), 3) if k == "config4" else (args.steps, args.warmup)
        r = measure_workload(ctx, k, st, wu, "samples" if k == "config5" else "tiles", with_clocks=False, parity_check=True)
        if r is not None:
            also[k] = {"metric": METRIC, "unit": "Mrays/s", **r}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sd = make_scene(key)
            cam, params = make_camera_params(key, sd)
            c = CpuArm(key, sd, cam, params).measure(budget_s=15.0)
            cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"metric": METRIC, "value": head["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",      # the same frame at every N
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        for k in ("config", "clocks", "e2e", "gpu_launches", "roofline"):
            line[k] = head[k]
        for k in ("combine_parity", "combine_max_rel_diff", "per_step_ms"):
            if k in head:
                line[k] = head[k]
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if also:
            line["also"] = also
        print(json.dumps(line))
    if world > 1:
        dist_t.barrier()
        ctx.comm.close()
        dist_t.destroy_process_group()
    return 0


def _json_only_stdout():
    """The contract is ONE JSON line on stdout. Native libraries write there too (NCCL prints "NCCL version ..." on the first communicator when
    NCCL_DEBUG is set, as it is on the GPU boxes), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the
    original descriptor."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    _json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--partition", default=None, choices=["samples", "tiles", "ranges"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--also", default=None, help="comma list of extra workloads measured in the same process ('none' to skip); default config2,config4 at N = 1, config4 at N > 1")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.also is None:
        args.also = ("config2,config4" if world == 1 else "config4") if args.workload == DEFAULT_WORKLOAD else "none"
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    rc = main()
    sys.stdout.flush()
    sys.exit(rc)
