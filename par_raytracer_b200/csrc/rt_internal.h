// rt_internal.h -- host-side declarations shared by the translation units of librt_b200.so:
//   rt_scene.cu     errors, scene upload, GPU hierarchy build, introspection
//   rt_render.cu    per-render pool, wave scheduler, rt_render* / rt_trace_* entry points
//   rt_loadtime.cu  tone map, BuildHierarchy over mesh groups, tangents, height -> normal map
//   rt_comm.cu      multi-GPU combine (NCCL / peer memory) that replaces MPI_Gather
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_types.cuh"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
extern thread_local std::string g_rt_err;
#define g_err g_rt_err

int rt_fail(int code, const char *fmt, ...);
#define fail rt_fail

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CKL(name)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                                              \
        if (e_ != cudaSuccess) return fail(RT_ERR_CUDA, "launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// device buffer bookkeeping
// ---------------------------------------------------------------------------------------------
// A caller's host buffer kept page-locked between calls (RT_FLAG_PIN_HOST)
struct PinnedHost {
    void *ptr = nullptr; size_t bytes = 0;
    void release() { if (ptr) { cudaHostUnregister(ptr); (void)cudaGetLastError(); ptr = nullptr; bytes = 0; } }
    void pin(void *p, size_t n) {
        if (p == ptr && n <= bytes) return;
        release();
        if (cudaHostRegister(p, n, cudaHostRegisterDefault) == cudaSuccess) { ptr = p; bytes = n; } else (void)cudaGetLastError();
    }
};

struct DevArena {
    std::vector<void *> ptrs;
    template <typename T> cudaError_t alloc(T **p, size_t n) {
        *p = nullptr;
        if (n == 0) n = 1;
        cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    void release() {
        for (void *p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

struct WaveTotals { unsigned long long closest, shadow, waves, pad; };

struct Pool {                 // per-render working set, kept between calls and grown on demand
    DevArena mem;
    uint32_t capacity = 0, depth = 0, lights = 0;
    PathPool paths{};
    RayQueue q[2]{};
    HitRec *hits = nullptr;
    uint32_t *ray_cnt = nullptr;       // per-path ray counts (PathPool::ray_cnt points here during the adaptive loop)
    ShadowQueue shadow{};
    uint32_t *counts = nullptr;        // [0],[1]: ray queue sizes (ping-pong), [2 + l]: shadow queue size of light l
    float4 *acc_extra = nullptr;       // per-path accumulators of lights >= 1
    uint32_t *next = nullptr;          // work counter of the wave trace kernel
    WaveTotals *totals = nullptr;
    TraceCounters *tcount = nullptr;
    uint32_t *h_counts = nullptr;      // pinned ring of count read-backs
    cudaEvent_t count_ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

struct rt_scene {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevArena mem;
    DevScene d{};
    std::vector<float> spec_intensity;   // per material (+ default), for the Phong-lobe table
    float4 *spec_dir = nullptr;
    uint32_t spec_dir_ss = 0;
    uint32_t n_lights = 0;
    uint64_t info[8] = {0};
    rt_stats stats{};
    Pool pool;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // grow-only scratch of render calls (no cudaMalloc / cudaFree -- an implicit device sync -- per frame)
    float4 *accum = nullptr; size_t accum_cap = 0;
    uint32_t *ids = nullptr; size_t ids_cap = 0;
    float *out_stage = nullptr; size_t out_stage_cap = 0;
    float4 *scratch = nullptr; size_t scratch_cap = 0;        // adaptive sampling: per-pixel sample colours
    uint32_t *ad_u32 = nullptr; size_t ad_u32_cap = 0;        // adaptive sampling: nsamples + 2 x (pixel, local) lists + counter
    uint32_t last_adaptive_pixels = 0;
    PinnedHost pinned_out;               // rt_render's output buffer (RT_FLAG_PIN_HOST)
    uint32_t pool_limit_cached = 0;      // path slots the per-render pool may grow to (decided at the first render from the free device memory)
    std::vector<cudaEvent_t> tev;        // per-wave kernel timing (RT_FLAG_TIME_KERNELS): 4 events per wave
    size_t tev_used = 0;
    std::vector<std::pair<uint64_t, uint64_t>> wave_log;   // (closest, shadow) rays per issued wave, aligned with tev (RT_B200_WAVE_LOG=1)
    int sm_count = 148;
    int bounds = RT_BOUNDS_QBOX;         // child bound of the traversal: quantised boxes (default), float boxes or sphere + slab (RT_B200_BOUNDS)
    int trace_grid = 148 * 8;            // persistent grid of k_trace_wave: resident blocks of the whole chip
    int logic_grid = 148 * 6;            // same for k_logic
};

// rt_scene.cu
bool rt_place_quant_grid(double lo, double hi, float *step_out, float *mid_out, double *base_out);
// rt_render.cu: persistent-grid sizes of the trace / shading kernels for this scene's child bound
int rt_render_configure(rt_scene *sc);
// rt_render.cu: rt_render_device with a device-resident, pre-validated pixel list (the cached tile partition of rt_comm.cu)
int rt_render_device_ids(rt_scene *scene, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height, const uint32_t *pixel_ids_device,
                         uint32_t pixel_count, uint32_t sample_begin, uint32_t sample_count, uint32_t flags, float *out_rgba_device, rt_counters *out_counters);
