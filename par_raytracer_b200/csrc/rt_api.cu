// rt_api.cu -- the C ABI of include/rt_b200.h: scene upload, GPU hierarchy build, wave scheduling.
// Host code in this file only moves data, sizes launches and builds small decode tables; every hit,
// ray and colour is computed by the kernels in rt_trace.cuh / rt_shade.cuh. There is no CPU fallback.
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
#include "rt_build.cuh"
#include "rt_groups.cuh"
#include "rt_preprocess.cuh"
#include "rt_common.cuh"
#include "rt_rng.cuh"
#include "rt_raygen.cuh"
#include "rt_shade.cuh"
#include "rt_trace.cuh"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

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

extern "C" const char *rt_last_error(void) { return g_err.c_str(); }
extern "C" int rt_abi_version(void) { return RT_ABI_VERSION; }

static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// device buffer bookkeeping
// ---------------------------------------------------------------------------------------------
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
    std::vector<cudaEvent_t> tev;        // per-wave kernel timing (RT_FLAG_TIME_KERNELS): 4 events per wave
    size_t tev_used = 0;
    std::vector<std::pair<uint64_t, uint64_t>> wave_log;   // (closest, shadow) rays per issued wave, aligned with tev (RT_B200_WAVE_LOG=1)
    int sm_count = 148;
    int bounds = RT_BOUNDS_QBOX;         // child bound of the traversal: quantised boxes (default), float boxes or sphere + slab (RT_B200_BOUNDS)
    int trace_grid = 148 * 8;            // persistent grid of k_trace_wave: resident blocks of the whole chip
    int logic_grid = 148 * 6;            // same for k_logic
};

// ---------------------------------------------------------------------------------------------
// host-built decode tables (the only libm calls of the product; same glibc the reference would use)
// ---------------------------------------------------------------------------------------------
#define RT_PI32 (3.1415927f)                                   // brt.h:23

static float host_radical_inverse(uint32_t bits) {             // raytracer.cpp:273-282
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)(bits * 2.3283064365386963e-10);
}

static void host_srgb_lut(float *lut) {                        // color.h:13-21 over texture.cpp:44-48's 256 inputs
    const float one_over_255 = 1.0f / 255.0f;
    for (int i = 0; i < 256; ++i) {
        volatile float srgb = (float)i * one_over_255;
        lut[i] = srgb <= 0.04045f ? srgb / 12.92f : powf((srgb + 0.055f) / 1.055f, 2.4f);
    }
}

static void host_hammersley_dirs(float4 *out) {                // raytracer.cpp:284-288 + 322-328 for i in [0, 1024)
    for (uint32_t i = 0; i < 1024; ++i) {
        volatile float xi_x = (float)i / (float)1024u;
        volatile float xi_y = host_radical_inverse(i);
        volatile float phi = xi_y * 2.0f * RT_PI32;
        volatile float cp = cosf(phi);
        volatile float sp = sinf(phi);
        volatile float ct = sqrtf(1.0f - xi_x);
        volatile float st = sqrtf(1.0f - ct * ct);
        out[i] = make_float4(cp * st, sp * st, ct, 0.0f);
    }
}

static void host_phong_dirs(const std::vector<float> &spec_intensity, uint32_t ss, std::vector<float4> &out) {   // raytracer.cpp:290-300
    out.resize(spec_intensity.size() * (size_t)std::max(1u, ss));
    for (size_t m = 0; m < spec_intensity.size(); ++m) {
        for (uint32_t s = 0; s < ss; ++s) {
            volatile float xi_x = (float)s / (float)ss;
            volatile float xi_y = host_radical_inverse(s);
            volatile float phi = 2.0f * RT_PI32 * xi_x;
            volatile float cp = cosf(phi);
            volatile float sp = sinf(phi);
            volatile float ct = powf(1.0f - xi_y, 1.0f / (spec_intensity[m] + 1.0f));
            volatile float st = sqrtf(1.0f - (ct * ct));
            out[m * ss + s] = make_float4(cp * st, sp * st, ct, 0.0f);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// quantisation grid of QNode (host): 32764 steps across [lo, hi] per axis, at least one step of margin below lo.
// The kernel decodes plane q as mid + (32768 + q) * step from these two FLOATS, so the grid is built from exactly them.
// A scene far from the origin relative to its size rounds `mid` coarsely: the step is widened until the grid covers [lo, hi].
// Returns false when no grid can (non-finite bounds).
// ---------------------------------------------------------------------------------------------
static bool place_quant_grid(double lo, double hi, float *step_out, float *mid_out, double *base_out) {
    float step = (float)std::max((hi - lo) / 32764.0, 1e-30);
    if (!(step > 0.0f)) step = 1e-30f;
    float mid = 0.0f; double base = 0.0;
    bool ok = false;
    for (int it = 0; it < 200 && !ok; ++it) {
        mid = (float)(lo - (double)step - 32768.0 * (double)step);
        base = (double)mid + 32768.0 * (double)step;
        ok = base <= lo && base + 32767.0 * (double)step >= hi;
        if (!ok) step *= 1.25f;
    }
    *step_out = step; *mid_out = mid; *base_out = base;
    return ok;
}

extern "C" int rt_quant_grid(const float lo[3], const float hi[3], float step[3], float mid[3], int *ok) {
    g_err.clear();
    if (!lo || !hi || !step || !mid || !ok) return fail(RT_ERR_ARG, "null argument");
    *ok = 1;
    for (int a = 0; a < 3; ++a) { double b; if (!place_quant_grid(lo[a], hi[a], &step[a], &mid[a], &b)) *ok = 0; }
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// hierarchy build orchestration (kernels in rt_build.cuh)
// ---------------------------------------------------------------------------------------------
static int build_hierarchy(rt_scene *sc, const BuildInput &bin, const GatherInput &gin, bool has_tangents) {
    cudaStream_t st = sc->stream;
    const uint32_t n = bin.n_tris;
    DevArena tmp;
    auto done = [&](int rc) { tmp.release(); return rc; };
#define CKB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
#define CKLB(name) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_))); } while (0)

    cudaEvent_t e0, e1;
    CKB(cudaEventCreate(&e0)); CKB(cudaEventCreate(&e1));
    CKB(cudaEventRecord(e0, st));

    // final arrays
    TriRec *tris; uint32_t *tri_rank, *tri_vertex0; int32_t *tri_object; float4 *tri_uv, *tri_nrm, *tri_tan = nullptr;
    CKB(sc->mem.alloc(&tris, n)); CKB(sc->mem.alloc(&tri_rank, n)); CKB(sc->mem.alloc(&tri_vertex0, n));
    CKB(sc->mem.alloc(&tri_object, n)); CKB(sc->mem.alloc(&tri_uv, 2 * (size_t)n)); CKB(sc->mem.alloc(&tri_nrm, 3 * (size_t)n));
    if (has_tangents) CKB(sc->mem.alloc(&tri_tan, 3 * (size_t)n));

    uint32_t n_pad = BITONIC_TILE;
    while (n_pad < n) n_pad <<= 1;
    float4 *tri_sphere, *tri_lo, *tri_hi, *tri_nrm0, *tri_slab; uint32_t *bounds; uint64_t *keys; uint32_t *vals;
    CKB(tmp.alloc(&tri_sphere, n)); CKB(tmp.alloc(&tri_lo, n)); CKB(tmp.alloc(&tri_hi, n)); CKB(tmp.alloc(&tri_nrm0, n)); CKB(tmp.alloc(&tri_slab, n));
    CKB(tmp.alloc(&bounds, 8)); CKB(tmp.alloc(&keys, n_pad)); CKB(tmp.alloc(&vals, n_pad));
    {
        uint32_t hb[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u};
        CKB(cudaMemcpyAsync(bounds, hb, sizeof(hb), cudaMemcpyHostToDevice, st));
    }
    k_tri_spheres<<<cdiv(n, 256), 256, 0, st>>>(bin, tri_sphere, tri_lo, tri_hi, tri_nrm0, tri_slab, bounds); CKLB("k_tri_spheres");
    k_morton<<<cdiv(n_pad, 256), 256, 0, st>>>(n, n_pad, tri_sphere, bounds, keys, vals); CKLB("k_morton");
    // bitonic sort
    k_bitonic_shared<<<n_pad / BITONIC_TILE, 1024, 0, st>>>(keys, vals, 2, BITONIC_TILE, 0); CKLB("k_bitonic_shared");
    for (uint64_t k = 2ull * BITONIC_TILE; k <= n_pad; k <<= 1) {
        for (uint32_t j = (uint32_t)(k >> 1); j >= BITONIC_TILE; j >>= 1) {
            k_bitonic_global<<<cdiv(n_pad, 256), 256, 0, st>>>(keys, vals, n_pad, j, (uint32_t)k); CKLB("k_bitonic_global");
        }
        k_bitonic_shared<<<n_pad / BITONIC_TILE, 1024, 0, st>>>(keys, vals, (uint32_t)k, (uint32_t)k, 1); CKLB("k_bitonic_shared");
    }

    // temp tree
    const uint32_t n_total = 2 * n - 1;
    TempTree t;
    CKB(tmp.alloc(&t.c0, n_total)); CKB(tmp.alloc(&t.c1, n_total)); CKB(tmp.alloc(&t.parent, n_total));
    CKB(tmp.alloc(&t.size, n_total)); CKB(tmp.alloc(&t.kept, n_total)); CKB(tmp.alloc(&t.sphere, n_total));
    CKB(tmp.alloc(&t.lo, n_total)); CKB(tmp.alloc(&t.hi, n_total)); CKB(tmp.alloc(&t.nsum, n_total)); CKB(tmp.alloc(&t.slab, n_total));
    int32_t *cn[2]; uint32_t *nn, *slot_tri; uint64_t *flags, *scan, *bsums, *total;
    CKB(tmp.alloc(&cn[0], n)); CKB(tmp.alloc(&cn[1], n)); CKB(tmp.alloc(&slot_tri, n));
    CKB(tmp.alloc(&nn, n)); CKB(tmp.alloc(&flags, n)); CKB(tmp.alloc(&scan, n));
    CKB(tmp.alloc(&bsums, cdiv(n, SCAN_TILE) + 1)); CKB(tmp.alloc(&total, 1));
    uint32_t *tri_offset, *kept_index, *max_depth, *rot_visit, *rot_count;
    CKB(tmp.alloc(&tri_offset, n_total)); CKB(tmp.alloc(&kept_index, n_total)); CKB(tmp.alloc(&max_depth, 1));
    CKB(tmp.alloc(&rot_visit, n_total)); CKB(tmp.alloc(&rot_count, 1)); CKB(cudaMemsetAsync(rot_count, 0, 4, st));

    uint32_t kept_nodes = 0, depth = 0, iterations = 0;
    int32_t root_temp = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        // attempt 1: strict (2k, 2k+1) pairing -> balanced tree of depth ceil(log2 n); mode 2 = surface-area search cost (default; RT_B200_PLOC_COST=diagonal: the squared-diagonal cost)
        const char *pc = getenv("RT_B200_PLOC_COST");
        const int pair_mode = (attempt || (pc && strcmp(pc, "pairs") == 0)) ? 1 : ((pc && strcmp(pc, "diagonal") == 0) ? 0 : 2);
        k_ploc_init<<<cdiv(n, 256), 256, 0, st>>>(n, vals, tri_sphere, tri_lo, tri_hi, tri_nrm0, tri_slab, cn[0], t); CKLB("k_ploc_init");
        uint32_t m = n, created = 0;
        int cur = 0;
        iterations = 0;
        while (m > 1) {
            k_ploc_nn<<<cdiv(m, 256), 256, 0, st>>>(m, cn[cur], t, nn, pair_mode); CKLB("k_ploc_nn");
            k_ploc_flags<<<cdiv(m, 256), 256, 0, st>>>(m, nn, flags); CKLB("k_ploc_flags");
            uint32_t nb = cdiv(m, SCAN_TILE);
            k_scan_reduce<<<nb, SCAN_BLOCK, 0, st>>>(flags, m, bsums); CKLB("k_scan_reduce");
            k_scan_blocksums<<<1, 1024, 0, st>>>(bsums, nb, total); CKLB("k_scan_blocksums");
            k_scan_apply<<<nb, SCAN_BLOCK, 0, st>>>(flags, m, bsums, scan); CKLB("k_scan_apply");
            k_ploc_merge<<<cdiv(m, 256), 256, 0, st>>>(m, n, created, nn, flags, scan, cn[cur], cn[cur ^ 1], t);
            CKLB("k_ploc_merge");
            uint64_t h_total = 0;
            CKB(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, st));
            CKB(cudaStreamSynchronize(st));
            uint32_t merges = (uint32_t)(h_total >> 32), valid = (uint32_t)(h_total & 0xffffffffull);
            if (merges == 0 || valid != m - merges) return done(fail(RT_ERR_STATE, "hierarchy build made no progress (m=%u merges=%u valid=%u)", m, merges, valid));
            created += merges;
            m = valid;
            cur ^= 1;
            iterations++;
        }
        CKB(cudaMemcpyAsync(&root_temp, cn[cur], 4, cudaMemcpyDeviceToHost, st));
        if (attempt == 0 && n > 2) {          // local restructuring sweeps (RT_B200_ROTATE; measured 0 / 1 / 2 / 4 sweeps at 10 M triangles: 22.6 / 22.4 / 22.3 / 22.3 ms)
            const char *re = getenv("RT_B200_ROTATE");
            const int sweeps = re ? std::max(0, std::min(8, atoi(re))) : 2;
            for (int sw = 0; sw < sweeps; ++sw) {
                CKB(cudaMemsetAsync(rot_visit, 0, (size_t)n_total * 4, st));
                k_rotate<<<cdiv(n, 256), 256, 0, st>>>(n, t, rot_visit, rot_count); CKLB("k_rotate");
            }
        }
        CKB(cudaMemsetAsync(max_depth, 0, 4, st));
        k_layout<<<cdiv(n_total, 256), 256, 0, st>>>(n_total, t, tri_offset, kept_index, max_depth); CKLB("k_layout");
        CKB(cudaMemcpyAsync(&depth, max_depth, 4, cudaMemcpyDeviceToHost, st));
        CKB(cudaMemcpyAsync(&kept_nodes, t.kept + (n_total - 1), 4, cudaMemcpyDeviceToHost, st));
        CKB(cudaStreamSynchronize(st));
        if (root_temp != (int32_t)(n_total - 1) && n > 1) return done(fail(RT_ERR_STATE, "hierarchy root mismatch"));
        if (depth + 2 <= RT_STACK_MAX) break;
        if (attempt == 1) return done(fail(RT_ERR_STATE, "hierarchy depth %u exceeds traversal stack", depth));
    }

    // quantisation grid of QNode: 32766 steps across the root box, one step of margin below it
    double qb[3] = {0, 0, 0}, qs[3] = {1, 1, 1};
    bool grid_ok = true;
    if (n >= 1) {
        float4 rlo, rhi;
        CKB(cudaMemcpyAsync(&rlo, t.lo + (n_total - 1), sizeof(rlo), cudaMemcpyDeviceToHost, st));
        CKB(cudaMemcpyAsync(&rhi, t.hi + (n_total - 1), sizeof(rhi), cudaMemcpyDeviceToHost, st));
        CKB(cudaStreamSynchronize(st));
        const float lo3[3] = {rlo.x, rlo.y, rlo.z}, hi3[3] = {rhi.x, rhi.y, rhi.z};
        for (int a = 0; a < 3; ++a) {
            if (!place_quant_grid(lo3[a], hi3[a], &sc->d.qstep[a], &sc->d.qmid[a], &qb[a])) grid_ok = false;     // non-finite extents
            qs[a] = (double)sc->d.qstep[a];
        }
    }
    if (!grid_ok && sc->bounds == RT_BOUNDS_QBOX) sc->bounds = RT_BOUNDS_BOX;   // float boxes need no grid (inf / NaN vertices behave as they do there)
    // only the node array of the selected child bound is built (RT_B200_BOUNDS)
    HNode *nodes = nullptr; BNode *bnodes = nullptr; QNode *qnodes = nullptr;
    if (sc->bounds == RT_BOUNDS_SPHERE) CKB(sc->mem.alloc(&nodes, std::max(1u, kept_nodes)));
    else if (sc->bounds == RT_BOUNDS_BOX) CKB(sc->mem.alloc(&bnodes, std::max(1u, kept_nodes)));
    else CKB(sc->mem.alloc(&qnodes, std::max(1u, kept_nodes)));
    if (n > 1) {
        k_slot_to_tri<<<cdiv(n, 256), 256, 0, st>>>(n, vals, tri_offset, slot_tri); CKLB("k_slot_to_tri");
        if (sc->bounds == RT_BOUNDS_SPHERE) { k_refit<<<cdiv((uint64_t)(n - 1) * 32, 256), 256, 0, st>>>(n, n_total, t, tri_offset, slot_tri, bin); CKLB("k_refit"); }
    }
    if (kept_nodes > 0) {
        k_emit_nodes<<<cdiv(n - 1, 256), 256, 0, st>>>(n, n_total, t, tri_offset, kept_index, nodes, bnodes, qnodes, qb[0], qb[1], qb[2], qs[0], qs[1], qs[2]); CKLB("k_emit_nodes");
        sc->d.root = 0;
    } else {
        sc->d.root = -(int)(1u + 0u * 8u + n);     // the whole scene is one cluster (n <= RT_LEAF_MAX)
    }
    k_gather<<<cdiv(n, 256), 256, 0, st>>>(gin, vals, tri_offset, tris, tri_rank, tri_uv, tri_nrm, tri_tan, tri_vertex0, tri_object);
    CKLB("k_gather");
    CKB(cudaEventRecord(e1, st));
    CKB(cudaStreamSynchronize(st));
    float ms = 0;
    CKB(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);

    sc->d.nodes = nodes; sc->d.bnodes = bnodes; sc->d.qnodes = qnodes; sc->d.tris = tris; sc->d.tri_rank = tri_rank; sc->d.tri_uv = tri_uv; sc->d.tri_nrm = tri_nrm;
    sc->d.tri_tan = tri_tan; sc->d.tri_vertex0 = tri_vertex0; sc->d.tri_object = tri_object;
    sc->d.n_tris = n; sc->d.n_nodes = kept_nodes;
    {   // every sphere lies inside the root sphere: |c|_1 + r <= |c_root|_1 + sqrt(3) * 2 r_root + r_root
        float4 rs;
        CKB(cudaMemcpy(&rs, t.sphere + (n_total - 1), sizeof(rs), cudaMemcpyDeviceToHost));
        sc->d.cull_bound = fabsf(rs.x) + fabsf(rs.y) + fabsf(rs.z) + 4.5f * rs.w;
    }
    sc->info[0] = n; sc->info[1] = 0; sc->info[2] = kept_nodes; sc->info[3] = depth;
    sc->info[4] = (uint64_t)kept_nodes * (sc->bounds == RT_BOUNDS_QBOX ? sizeof(QNode) : sc->bounds == RT_BOUNDS_BOX ? sizeof(BNode) : sizeof(HNode)); sc->info[5] = (uint64_t)n * sizeof(TriRec);
    sc->info[6] = (uint64_t)(ms * 1000.0f); sc->info[7] = iterations;
    return done(RT_OK);
#undef CKB
#undef CKLB
}

// ---------------------------------------------------------------------------------------------
// rt_scene_create / destroy
// ---------------------------------------------------------------------------------------------
template <typename T> static cudaError_t upload(DevArena &a, cudaStream_t st, T **dst, const T *src, size_t n) {
    cudaError_t e = a.alloc(dst, n);
    if (e != cudaSuccess) return e;
    if (n && src) e = cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}

extern "C" void rt_scene_destroy(rt_scene *sc) {
    if (!sc) return;
    cudaSetDevice(sc->device);
    if (sc->stream) cudaStreamSynchronize(sc->stream);
    sc->pool.mem.release();
    if (sc->accum) cudaFree(sc->accum);
    if (sc->ids) cudaFree(sc->ids);
    if (sc->out_stage) cudaFree(sc->out_stage);
    if (sc->scratch) cudaFree(sc->scratch);
    if (sc->ad_u32) cudaFree(sc->ad_u32);
    if (sc->pool.h_counts) cudaFreeHost(sc->pool.h_counts);
    for (int k = 0; k < 4; ++k) if (sc->pool.count_ev[k]) cudaEventDestroy(sc->pool.count_ev[k]);
    sc->mem.release();
    if (sc->ev0) cudaEventDestroy(sc->ev0);
    if (sc->ev1) cudaEventDestroy(sc->ev1);
    for (cudaEvent_t e : sc->tev) cudaEventDestroy(e);
    if (sc->stream) cudaStreamDestroy(sc->stream);
    delete sc;
}

extern "C" int rt_scene_create(const rt_scene_desc *desc, int device, rt_scene **out_scene) {
    g_err.clear();
    if (!desc || !out_scene) return fail(RT_ERR_ARG, "null argument");
    *out_scene = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(RT_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    const uint32_t G = desc->n_groups;
    if (G && (!desc->group_first || !desc->idx_positions || !desc->idx_texcoords || !desc->idx_normals || !desc->group_material))
        return fail(RT_ERR_ARG, "group arrays missing");
    const uint64_t n_idx = G ? desc->group_first[G] : 0;
    if (n_idx % 3) return fail(RT_ERR_ARG, "index count %llu not a multiple of 3", (unsigned long long)n_idx);
    if (n_idx / 3 > 200000000ull) return fail(RT_ERR_ARG, "too many triangles");
    const uint32_t n_tris = (uint32_t)(n_idx / 3);
    for (uint32_t g = 0; g < G; ++g)
        if (desc->group_first[g + 1] < desc->group_first[g] || (desc->group_first[g + 1] - desc->group_first[g]) % 3)
            return fail(RT_ERR_ARG, "group %u index range invalid", g);
    if (n_tris && (!desc->positions || !desc->texcoords || !desc->normals)) return fail(RT_ERR_ARG, "vertex streams missing");
    for (uint64_t i = 0; i < n_idx; ++i) {
        if (desc->idx_positions[i] >= desc->n_positions || desc->idx_texcoords[i] >= desc->n_texcoords ||
            desc->idx_normals[i] >= desc->n_normals)
            return fail(RT_ERR_ARG, "vertex index out of range at %llu", (unsigned long long)i);
    }
    CK(cudaSetDevice(device));
    rt_scene *sc = new rt_scene;
    sc->device = device;
    auto bail = [&](int rc) { rt_scene_destroy(sc); return rc; };
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    CKS(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
    CKS(cudaEventCreate(&sc->ev0)); CKS(cudaEventCreate(&sc->ev1));
    CKS(cudaDeviceGetAttribute(&sc->sm_count, cudaDevAttrMultiProcessorCount, device));
    {
        int per_sm = 0;
        const char *be = getenv("RT_B200_BOUNDS");      // "box" / "sphere": the float-box and sphere + slab child bounds (kept for the comparison in profiles/)
        sc->bounds = (be && strcmp(be, "sphere") == 0) ? RT_BOUNDS_SPHERE : (be && strcmp(be, "box") == 0) ? RT_BOUNDS_BOX : RT_BOUNDS_QBOX;
        if (sc->bounds == RT_BOUNDS_QBOX) CKS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wave<false, RT_BOUNDS_QBOX>, RT_TRACE_BLOCK, 0));
        else if (sc->bounds == RT_BOUNDS_BOX) CKS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wave<false, RT_BOUNDS_BOX>, RT_TRACE_BLOCK, 0));
        else CKS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wave<false, RT_BOUNDS_SPHERE>, RT_TRACE_BLOCK, 0));
        sc->trace_grid = sc->sm_count * std::max(1, per_sm);
        CKS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_logic, 128, 0));
        sc->logic_grid = sc->sm_count * std::max(1, per_sm) * 2;
    }
    cudaStream_t st = sc->stream;

    // ---- tie-break ranks: the reference's leaf encounter order (raytracer.cpp:168-172, 208-209) ----
    std::vector<uint32_t> rank_base(G ? G : 1, 0);
    std::vector<int32_t> group_object(G ? G : 1, -1);
    {
        std::vector<char> seen(G ? G : 1, 0);
        uint32_t running = 0;
        if (desc->n_spheres && desc->spheres && desc->sphere_group) {
            std::vector<uint32_t> stack; stack.push_back(0);
            uint64_t guard = 0;
            while (!stack.empty()) {
                uint32_t i = stack.back(); stack.pop_back();
                if (i >= desc->n_spheres || ++guard > 4ull * desc->n_spheres + 8) return bail(fail(RT_ERR_ARG, "malformed sphere hierarchy"));
                const rt_bsphere &s = desc->spheres[i];
                if (s.c0 && s.c1) { stack.push_back(s.c0); stack.push_back(s.c1); }
                else {
                    int32_t g = desc->sphere_group[i];
                    if (g < 0 || (uint32_t)g >= G || seen[g]) return bail(fail(RT_ERR_ARG, "sphere %u: bad mesh group %d", i, g));
                    seen[g] = 1; rank_base[g] = running; group_object[g] = (int32_t)i;
                    running += (desc->group_first[g + 1] - desc->group_first[g]) / 3;
                }
            }
            for (uint32_t g = 0; g < G; ++g) if (!seen[g]) return bail(fail(RT_ERR_ARG, "mesh group %u is in no leaf sphere", g));
        } else {
            for (uint32_t g = 0; g < G; ++g) { rank_base[g] = running; group_object[g] = (int32_t)g; running += (desc->group_first[g + 1] - desc->group_first[g]) / 3; }
        }
    }
    std::vector<int32_t> group_mat(G ? G : 1, 0);
    for (uint32_t g = 0; g < G; ++g) {
        int32_t m = desc->group_material[g];
        if (m >= (int32_t)desc->n_materials) return bail(fail(RT_ERR_ARG, "group %u material %d out of range", g, m));
        group_mat[g] = m < 0 ? (int32_t)desc->n_materials : m;
    }

    // ---- materials / textures / lights ----
    std::vector<DevMaterial> mats(desc->n_materials + 1);
    sc->spec_intensity.resize(desc->n_materials + 1);
    bool any_bump = false;
    for (uint32_t i = 0; i <= desc->n_materials; ++i) {
        const rt_material &m = i < desc->n_materials ? desc->materials[i] : desc->default_material;
        DevMaterial &d = mats[i];
        memset(&d, 0, sizeof(d));
        d.specular_intensity = m.specular_intensity; d.index_of_refraction = m.index_of_refraction; d.alpha = m.alpha;
        for (int k = 0; k < 3; ++k) { d.ambient[k] = m.ambient_color[k]; d.diffuse[k] = m.diffuse_color[k]; d.specular[k] = m.specular_color[k]; }
        d.tex_ambient = m.ambient_texture; d.tex_diffuse = m.diffuse_texture; d.tex_specular = m.specular_texture;
        d.tex_alpha = m.alpha_texture; d.tex_bump = m.bump_texture;
        const int32_t *tx[5] = {&d.tex_ambient, &d.tex_diffuse, &d.tex_specular, &d.tex_alpha, &d.tex_bump};
        for (int k = 0; k < 5; ++k) if (*tx[k] >= (int32_t)desc->n_textures) return bail(fail(RT_ERR_ARG, "material %u texture index out of range", i));
        if (d.tex_bump >= 0) any_bump = true;
        sc->spec_intensity[i] = m.specular_intensity;
    }
    if (any_bump && !desc->tangents) return bail(fail(RT_ERR_ARG, "bump-mapped material but no tangents"));
    std::vector<DevTexture> texs(desc->n_textures ? desc->n_textures : 1);
    std::vector<uint8_t> blob;
    for (uint32_t i = 0; i < desc->n_textures; ++i) {
        const rt_texture &t = desc->textures[i];
        if (!t.texels || t.channels < 1 || t.channels > 4 || t.size_x < 2 || t.size_y < 2) return bail(fail(RT_ERR_ARG, "texture %u invalid", i));
        texs[i].size_x = t.size_x; texs[i].size_y = t.size_y; texs[i].channels = t.channels; texs[i].offset = (uint32_t)blob.size();
        size_t bytes = (size_t)t.size_x * t.size_y * t.channels;
        if (blob.size() + bytes > 0xFFFFFFFFull) return bail(fail(RT_ERR_ARG, "textures exceed 4 GiB"));
        blob.insert(blob.end(), t.texels, t.texels + bytes);
    }
    std::vector<DevLight> lights(desc->n_lights ? desc->n_lights : 1);
    for (uint32_t i = 0; i < desc->n_lights; ++i) {
        const rt_light &l = desc->lights[i];
        DevLight &d = lights[i];
        memset(&d, 0, sizeof(d));
        if (l.type != RT_LIGHT_DIRECTIONAL && l.type != RT_LIGHT_POINT) return bail(fail(RT_ERR_ARG, "light %u: unrecognised type %d", i, l.type));
        d.type = l.type; d.falloff = l.falloff;
        for (int k = 0; k < 3; ++k) { d.color[k] = l.color[k]; d.position[k] = l.position[k]; d.facing[k] = l.facing[k]; }
    }
    sc->n_lights = desc->n_lights;

    DevMaterial *d_mats; DevTexture *d_texs; uint8_t *d_blob; DevLight *d_lights; float *d_lut; float4 *d_hamm;
    CKS(upload(sc->mem, st, &d_mats, mats.data(), mats.size()));
    CKS(upload(sc->mem, st, &d_texs, texs.data(), texs.size()));
    CKS(upload(sc->mem, st, &d_blob, blob.data(), blob.size()));
    CKS(upload(sc->mem, st, &d_lights, lights.data(), lights.size()));
    float lut[256]; host_srgb_lut(lut);
    std::vector<float4> hamm(1024); host_hammersley_dirs(hamm.data());
    CKS(upload(sc->mem, st, &d_lut, lut, 256));
    CKS(upload(sc->mem, st, &d_hamm, hamm.data(), 1024));
    sc->d.materials = d_mats; sc->d.textures = d_texs; sc->d.texels = d_blob; sc->d.lights = d_lights;
    sc->d.srgb_lut = d_lut; sc->d.hamm_dir = d_hamm; sc->d.spec_dir = nullptr;
    sc->d.n_materials = desc->n_materials; sc->d.n_lights = desc->n_lights;
    sc->stats.h2d_bytes = mats.size() * sizeof(DevMaterial) + blob.size();

    // ---- geometry: upload the reference's arrays as they are, build on the GPU ----
    sc->d.n_tris = 0; sc->d.root = 0;
    if (n_tris) {
        DevArena in;     // input arrays are only needed during the build
        auto bail2 = [&](int rc) { in.release(); return bail(rc); };
#define CKI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail2(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
        float *d_pos, *d_tc, *d_nrm, *d_tan = nullptr; uint32_t *d_ip, *d_it, *d_in, *d_gf, *d_rb; int32_t *d_go, *d_gm;
        CKI(upload(in, st, &d_pos, desc->positions, 3 * (size_t)desc->n_positions));
        CKI(upload(in, st, &d_tc, desc->texcoords, 2 * (size_t)desc->n_texcoords));
        CKI(upload(in, st, &d_nrm, desc->normals, 3 * (size_t)desc->n_normals));
        if (any_bump) CKI(upload(in, st, &d_tan, desc->tangents, 3 * (size_t)desc->n_normals));
        CKI(upload(in, st, &d_ip, desc->idx_positions, (size_t)n_idx));
        CKI(upload(in, st, &d_it, desc->idx_texcoords, (size_t)n_idx));
        CKI(upload(in, st, &d_in, desc->idx_normals, (size_t)n_idx));
        CKI(upload(in, st, &d_gf, desc->group_first, (size_t)G + 1));
        CKI(upload(in, st, &d_rb, rank_base.data(), (size_t)G));
        CKI(upload(in, st, &d_go, group_object.data(), (size_t)G));
        CKI(upload(in, st, &d_gm, group_mat.data(), (size_t)G));
        sc->stats.h2d_bytes += 4ull * (3ull * desc->n_positions + 2ull * desc->n_texcoords + 3ull * desc->n_normals * (any_bump ? 2 : 1) + 3ull * n_idx);
        BuildInput bin; bin.positions = d_pos; bin.idx_positions = d_ip; bin.group_first = d_gf; bin.n_groups = G; bin.n_tris = n_tris;
        GatherInput gin; gin.positions = d_pos; gin.texcoords = d_tc; gin.normals = d_nrm; gin.tangents = d_tan;
        gin.idx_p = d_ip; gin.idx_t = d_it; gin.idx_n = d_in; gin.group_first = d_gf; gin.group_rank_base = d_rb;
        gin.group_object = d_go; gin.group_material = d_gm; gin.n_groups = G; gin.n_tris = n_tris;
        int rc = build_hierarchy(sc, bin, gin, any_bump);
        if (rc != RT_OK) return bail2(rc);
        in.release();
#undef CKI
    }
    CKS(cudaStreamSynchronize(st));
#undef CKS
    *out_scene = sc;
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// pool
// ---------------------------------------------------------------------------------------------
static uint32_t pool_limit() {
    const char *e = getenv("RT_B200_POOL");
    uint32_t v = e ? (uint32_t)strtoul(e, nullptr, 10) : 0;
    return v ? v : (1u << 25);
}

static int ensure_pool(rt_scene *sc, uint32_t capacity, uint32_t depth) {
    Pool &p = sc->pool;
    uint32_t lights = std::max(1u, sc->n_lights);
    depth = std::max(1u, depth);
    if (p.capacity >= capacity && p.depth >= depth && p.lights >= lights) return RT_OK;
    CK(cudaStreamSynchronize(sc->stream));
    p.mem.release();
    capacity = std::max(capacity, p.capacity); depth = std::max(depth, p.depth);
    p.capacity = p.depth = 0;
    size_t c = capacity;
    CK(p.mem.alloc(&p.paths.rng_cx, c)); CK(p.mem.alloc(&p.paths.rng_seed, c));
    CK(p.mem.alloc(&p.paths.acc, c)); CK(p.mem.alloc(&p.paths.node_T, c));
    CK(p.mem.alloc(&p.paths.frames, c * RT_FRAME_F4 * depth));
    CK(p.mem.alloc(&p.ray_cnt, c));
    p.paths.ray_cnt = nullptr;           // switched on only by the adaptive loop
    p.paths.capacity = capacity;
    for (int k = 0; k < 2; ++k) { CK(p.mem.alloc(&p.q[k].o, c)); CK(p.mem.alloc(&p.q[k].d, c)); }
    CK(p.mem.alloc(&p.hits, c));
    CK(p.mem.alloc(&p.shadow.o, c * lights)); CK(p.mem.alloc(&p.shadow.rad, c * lights));
    CK(p.mem.alloc(&p.counts, 2 + lights)); CK(p.mem.alloc(&p.totals, 1)); CK(p.mem.alloc(&p.tcount, 1));
    CK(p.mem.alloc(&p.acc_extra, c * (lights - 1))); CK(p.mem.alloc(&p.next, 1));
    CK(cudaMemsetAsync(p.acc_extra, 0, std::max<size_t>(1, c * (lights - 1)) * sizeof(float4), sc->stream));
    p.shadow.count = p.counts + 2;
    p.shadow.capacity = capacity;
    if (p.h_counts) { cudaFreeHost(p.h_counts); p.h_counts = nullptr; }
    CK(cudaMallocHost((void **)&p.h_counts, 4 * (2 + lights) * sizeof(uint32_t)));
    for (int k = 0; k < 4; ++k) if (!p.count_ev[k]) CK(cudaEventCreateWithFlags(&p.count_ev[k], cudaEventDisableTiming));
    p.capacity = capacity; p.depth = depth; p.lights = lights;
    return RT_OK;
}

static int ensure_spec_table(rt_scene *sc, uint32_t ss) {
    if (sc->spec_dir && sc->spec_dir_ss == ss) return RT_OK;
    std::vector<float4> tab;
    host_phong_dirs(sc->spec_intensity, ss, tab);
    CK(cudaStreamSynchronize(sc->stream));
    float4 *d;
    CK(sc->mem.alloc(&d, tab.size()));
    CK(cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(float4), cudaMemcpyHostToDevice, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    sc->spec_dir = d; sc->spec_dir_ss = ss; sc->d.spec_dir = d;
    return RT_OK;
}

template <typename T> static int grow(rt_scene *sc, T **buf, size_t *cap, size_t need) {
    if (*cap >= need && *buf) return RT_OK;
    CK(cudaStreamSynchronize(sc->stream));
    if (*buf) CK(cudaFree(*buf));
    *buf = nullptr; *cap = 0;
    size_t n = std::max<size_t>(need, 1);
    CK(cudaMalloc((void **)buf, n * sizeof(T)));
    *cap = n;
    return RT_OK;
}

static DevParams to_dev_params(const rt_params *p) {
    DevParams d;
    d.ray_bias = p->ray_bias; d.reflection_samples = p->reflection_samples; d.spec_samples = p->spec_samples;
    d.bounce_depth = p->bounce_depth; d.bg[0] = p->background_color[0]; d.bg[1] = p->background_color[1];
    d.bg[2] = p->background_color[2]; d.pad = 0; d.base_seed = p->base_seed;
    return d;
}

static int check_params(const rt_params *p) {
    if (!p) return fail(RT_ERR_ARG, "params is null");
    if (p->bounce_depth > 200) return fail(RT_ERR_ARG, "bounce_depth %u > 200", p->bounce_depth);
    if (p->reflection_samples + p->spec_samples > 1000000u) return fail(RT_ERR_ARG, "too many reflection samples");
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// the wave loop: runs every live path of the pool to completion
// ---------------------------------------------------------------------------------------------
struct WaveCfg { bool count; };

static int wave_event(rt_scene *sc, bool on) {
    if (!on) return RT_OK;
    if (sc->tev_used == sc->tev.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        sc->tev.push_back(e);
    }
    CK(cudaEventRecord(sc->tev[sc->tev_used++], sc->stream));
    return RT_OK;
}

// sums the per-wave event intervals recorded since tev_used was reset (stream must be idle)
static int collect_wave_times(rt_scene *sc) {
    const bool log = getenv("RT_B200_WAVE_LOG") != nullptr;
    for (size_t i = 0; i + 3 < sc->tev_used; i += 4) {
        float a = 0, b = 0, c = 0;
        CK(cudaEventElapsedTime(&a, sc->tev[i], sc->tev[i + 1]));
        CK(cudaEventElapsedTime(&b, sc->tev[i + 1], sc->tev[i + 2]));
        CK(cudaEventElapsedTime(&c, sc->tev[i + 2], sc->tev[i + 3]));
        sc->stats.trace_ms += a; sc->stats.logic_ms += b; sc->stats.shadow_ms += c;
        if (log && i / 4 < sc->wave_log.size()) {
            auto &wl = sc->wave_log[i / 4];
            fprintf(stderr, "[wave %3zu] closest %9llu shadow %9llu  trace %7.3f ms (%6.0f Mrays/s)  logic %7.3f ms\n", i / 4, (unsigned long long)wl.first,
                    (unsigned long long)wl.second, a, a > 0 ? (wl.first + wl.second) / a / 1e3 : 0.0, b);
        }
    }
    sc->tev_used = 0;
    sc->wave_log.clear();
    return RT_OK;
}

static uint32_t env_knob(const char *name, uint32_t dflt) {
    const char *e = getenv(name);
    uint32_t v = e ? (uint32_t)atoi(e) : dflt;
    return (v < 1 || v > 32) ? dflt : v;
}
static void set_fetch_knobs(WaveQueues &w) {
    uint32_t a = 0, b = 0, c = 0, d = 0;      // read per launch (microseconds): lets one process sweep the knobs
    { a = env_knob("RT_B200_FETCH_MIN", RT_FETCH_MIN); b = env_knob("RT_B200_FETCH_PRIMARY", RT_FETCH_MIN); c = env_knob("RT_B200_FETCH_SHADOW", RT_FETCH_MIN);
              d = env_knob("RT_B200_LEAF_WAIT", RT_LEAF_WAIT); }
    w.fetch_min = a; w.fetch_min_primary = b; w.fetch_min_shadow = c; w.leaf_wait = d;
}

static WaveQueues wave_queues(rt_scene *sc, int cur, uint32_t n_closest_max) {
    Pool &p = sc->pool;
    WaveQueues w;
    w.closest = p.q[cur]; w.n_closest = p.counts + cur; w.closest_max = n_closest_max; w.hits = p.hits;
    w.shadow_o = p.shadow.o; w.shadow_dir = nullptr; w.rad = p.shadow.rad; w.n_shadow = p.shadow.count; w.shadow_stride = p.shadow.capacity;
    w.n_lights = sc->n_lights; w.acc = p.paths.acc; w.acc_extra = p.acc_extra; w.next = p.next;
    set_fetch_knobs(w);
    return w;
}

static PrimaryGen no_gen() { PrimaryGen g; memset(&g, 0, sizeof(g)); return g; }

static int launch_trace_wave(rt_scene *sc, float bias, const WaveQueues &w, const PrimaryGen &gen, uint64_t work_bound, bool count, TraceCounters *tc) {
    cudaStream_t st = sc->stream;
    CK(cudaMemsetAsync(w.next, 0, 4, st));
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sc->trace_grid, std::max<uint64_t>(1, (work_bound + RT_TRACE_BLOCK - 1) / RT_TRACE_BLOCK));
#define RT_TRACE_LAUNCH(B) do { if (count) k_trace_wave<true, B><<<grid, RT_TRACE_BLOCK, 0, st>>>(sc->d, bias, w, gen, tc); \
                                else k_trace_wave<false, B><<<grid, RT_TRACE_BLOCK, 0, st>>>(sc->d, bias, w, gen, tc); } while (0)
    if (sc->bounds == RT_BOUNDS_QBOX) RT_TRACE_LAUNCH(RT_BOUNDS_QBOX);
    else if (sc->bounds == RT_BOUNDS_BOX) RT_TRACE_LAUNCH(RT_BOUNDS_BOX);
    else RT_TRACE_LAUNCH(RT_BOUNDS_SPHERE);
#undef RT_TRACE_LAUNCH
    CKL("k_trace_wave");
    return RT_OK;
}

// Runs every live path of the pool to completion. Wave w: ONE trace launch (the pending nodes' closest-hit rays +
// the shadow rays the previous shading step queued), then the shading / bounce-generation step.
//
// The host never stalls the GPU: queue sizes live in device memory (kernels read them there), and wave w + 1 is
// enqueued -- with launch bounds taken from the newest counts the host already has -- BEFORE the host waits for
// wave w's 16-byte count read-back. The read-backs only decide when to stop and feed the statistics.
#define RT_COUNT_RING 4
static int run_waves(rt_scene *sc, const DevParams &prm, uint32_t n_first, uint32_t flags, uint64_t *launches, const PrimaryGen *first_gen = nullptr) {
    const bool timed = (flags & RT_FLAG_TIME_KERNELS) != 0;
    Pool &p = sc->pool;
    cudaStream_t st = sc->stream;
    const bool count = (flags & RT_FLAG_COUNTERS) != 0;
    const uint32_t L = sc->n_lights;
    const uint32_t stride = 2 + L;                      // words per read-back slot
    if (L) CK(cudaMemsetAsync(p.counts + 2, 0, 4 * L, st));

    uint32_t bound = n_first;                           // upper bound of the closest-hit queue of the wave being issued
    uint32_t in_closest[RT_COUNT_RING] = {0};           // what the host knew when it issued wave w (for the statistics)
    auto issue = [&](uint32_t w) -> int {
        const int cur = (int)(w & 1u);
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        const PrimaryGen gen = (w == 0 && first_gen) ? *first_gen : no_gen();     // wave 0 of a render: rays are generated in place
        { int rc_ = launch_trace_wave(sc, prm.ray_bias, wave_queues(sc, cur, bound), gen, (uint64_t)bound * (1 + L), count, p.tcount); if (rc_) return rc_; }
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        CK(cudaMemsetAsync(p.counts + (cur ^ 1), 0, 4, st));
        if (L) CK(cudaMemsetAsync(p.counts + 2, 0, 4 * L, st));
        k_logic<<<std::min(cdiv(std::max(1u, bound), 128), (uint32_t)sc->logic_grid), 128, 0, st>>>(sc->d, prm, p.paths, p.q[cur], p.hits, p.counts + cur, bound, p.q[cur ^ 1], p.counts + (cur ^ 1), p.shadow, gen);
        CKL("k_logic");
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        CK(cudaMemcpyAsync(p.h_counts + (size_t)(w % RT_COUNT_RING) * stride, p.counts, 4 * stride, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(p.count_ev[w % RT_COUNT_RING], st));
        *launches += 2;
        return RT_OK;
    };

    uint64_t sh_in = 0;                                 // shadow rays traced by the wave whose read-back we wait for
    uint32_t c_in = n_first;                            // closest rays traced by that wave
    { int rc_ = issue(0); if (rc_) return rc_; }
    for (uint32_t w = 0;; ++w) {
        { int rc_ = issue(w + 1); if (rc_) return rc_; }                       // run ahead by one wave
        CK(cudaEventSynchronize(p.count_ev[w % RT_COUNT_RING]));
        const uint32_t *hc = p.h_counts + (size_t)(w % RT_COUNT_RING) * stride;
        const int cur = (int)(w & 1u);
        uint32_t c_out = hc[cur ^ 1];
        uint64_t sh_out = 0;
        for (uint32_t l = 0; l < L; ++l) sh_out += hc[2 + l];
        sc->stats.closest_rays += c_in;
        sc->stats.shadow_rays += sh_in;
        sc->stats.waves += 1;
        if (timed) sc->wave_log.push_back(std::make_pair((uint64_t)c_in, sh_in));
        c_in = c_out; sh_in = sh_out;
        bound = c_out;                                  // queue sizes never grow: every live path emits at most one ray per wave
        if (c_out == 0 && sh_out == 0) break;           // the wave already in flight finds empty queues and does nothing
    }
    (void)in_closest;
    if (timed) sc->wave_log.push_back(std::make_pair((uint64_t)0, (uint64_t)0));      // the run-ahead wave that found empty queues
    if (L > 1) {
        k_fold_light_acc<<<cdiv(n_first, 256), 256, 0, st>>>(p.paths.acc, p.acc_extra, n_first, p.shadow.capacity, L - 1);
        CKL("k_fold_light_acc");
        *launches += 1;
    }
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// rt_render_device / rt_render
// ---------------------------------------------------------------------------------------------
static DevCamera to_dev_camera(const rt_camera *c) {
    DevCamera d;
    static_assert(sizeof(DevCamera) == sizeof(rt_camera), "camera layout");
    memcpy(&d, c, sizeof(d));
    return d;
}

static int render_impl(rt_scene *sc, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                       const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                       uint32_t sample_count, uint32_t flags, float *out_dev, cudaStream_t user_stream, rt_counters *out_counters) {
    if (!sc || !cam || !out_dev) return fail(RT_ERR_ARG, "null argument");
    int rc = check_params(params);
    if (rc) return rc;
    if (!width || !height) return fail(RT_ERR_ARG, "empty frame");
    const uint64_t frame = (uint64_t)width * height;
    if (!pixel_ids && (uint64_t)pixel_begin + pixel_count > frame) return fail(RT_ERR_ARG, "pixel range exceeds the frame");
    if (pixel_ids) for (uint32_t k = 0; k < pixel_count; ++k) if (pixel_ids[k] >= frame) return fail(RT_ERR_ARG, "pixel id %u out of range", pixel_ids[k]);
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    if (user_stream) {      // order after the caller's stream
        CK(cudaEventRecord(sc->ev1, user_stream));
        CK(cudaStreamWaitEvent(st, sc->ev1, 0));
    }
    if ((flags & RT_FLAG_ADAPTIVE) && params->min_samples < params->max_samples) {
        if (params->min_samples == 0) return fail(RT_ERR_ARG, "adaptive sampling needs min_samples >= 1");
        if (flags & RT_OUT_SUM) return fail(RT_ERR_ARG, "adaptive sampling resolves per pixel; RT_OUT_SUM is not meaningful");
        sample_count = params->min_samples;        // first loop of RenderPixel (main.cpp:237-243); the second follows below
    }
    memset(&sc->stats, 0, sizeof(sc->stats));
    sc->tev_used = 0;
    uint64_t launches = 0;
    if (pixel_count == 0 || sample_count == 0) { if (out_counters) memset(out_counters, 0, sizeof(*out_counters)); return RT_OK; }

    DevParams prm = to_dev_params(params);
    DevCamera dcam = to_dev_camera(cam);
    rc = ensure_spec_table(sc, params->spec_samples);
    if (rc) return rc;
    const uint32_t limit = pool_limit();
    const uint32_t spp_chunk = std::min(sample_count, limit);
    const uint32_t pix_per_batch = std::max(1u, std::min(pixel_count, limit / spp_chunk));
    uint64_t pool_want = (uint64_t)pix_per_batch * spp_chunk;
    if ((flags & RT_FLAG_ADAPTIVE) && params->min_samples < params->max_samples)      // room for the second loop's chunks of up to 32 samples per pixel
        pool_want = std::max<uint64_t>(pool_want, (uint64_t)pixel_count * std::min(32u, params->max_samples - params->min_samples));
    rc = ensure_pool(sc, (uint32_t)std::min<uint64_t>(pool_want, (uint64_t)limit), params->bounce_depth);
    if (rc) return rc;
    Pool &p = sc->pool;

    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    uint32_t *d_ids = nullptr;
    rc = grow(sc, &sc->accum, &sc->accum_cap, pixel_count);
    if (rc) return done(rc);
    float4 *accum = sc->accum;
    CKR(cudaMemsetAsync(accum, 0, (size_t)pixel_count * sizeof(float4), st));
    if (pixel_ids) {
        rc = grow(sc, &sc->ids, &sc->ids_cap, pixel_count);
        if (rc) return done(rc);
        d_ids = sc->ids;
        CKR(cudaMemcpyAsync(d_ids, pixel_ids, (size_t)pixel_count * 4, cudaMemcpyHostToDevice, st));
        sc->stats.h2d_bytes += (uint64_t)pixel_count * 4;
    }
    CKR(cudaMemsetAsync(p.totals, 0, sizeof(WaveTotals), st));
    CKR(cudaMemsetAsync(p.tcount, 0, sizeof(TraceCounters), st));
    CKR(cudaEventRecord(sc->ev0, st));

    const bool adaptive = (flags & RT_FLAG_ADAPTIVE) && params->min_samples < params->max_samples;
    const uint32_t max_s = params->max_samples;
    bool adaptive_counts = false;                     // ray_count = first pass + the rays of the ACCEPTED samples of the second loop
    unsigned long long adaptive_first_pass_rays = 0, adaptive_accepted_rays = 0;
    uint32_t *d_nsamples = nullptr;
    if (adaptive) {
        rc = grow(sc, &sc->scratch, &sc->scratch_cap, (size_t)pixel_count * max_s);
        if (rc) return done(rc);
        rc = grow(sc, &sc->ad_u32, &sc->ad_u32_cap, 5 * (size_t)pixel_count + 8);
        if (rc) return done(rc);
        d_nsamples = sc->ad_u32;
    }
    for (uint32_t p0 = 0; p0 < pixel_count; p0 += pix_per_batch) {
        const uint32_t npix = std::min(pix_per_batch, pixel_count - p0);
        for (uint32_t s0 = 0; s0 < sample_count; s0 += spp_chunk) {
            const uint32_t ns = std::min(spp_chunk, sample_count - s0);
            const uint32_t n_slots = npix * ns;
            PrimaryGen gen;
            memset(&gen, 0, sizeof(gen));
            gen.cam = dcam; gen.base_seed = prm.base_seed; gen.pixel_ids = d_ids; gen.n_slots = n_slots; gen.spp = ns; gen.width = width;
            gen.pixel_begin = pixel_begin; gen.pixel_local0 = p0; gen.sample_begin = sample_begin + s0; gen.jitter_scale = 0.5f; gen.enabled = 1;
            rc = run_waves(sc, prm, n_slots, flags, &launches, &gen);
            if (rc) return done(rc);
            if (adaptive) k_resolve_scratch<<<cdiv(npix, 128), 128, 0, st>>>(p.paths.acc, npix, ns, accum, sc->scratch, pixel_count, p0, s0);
            else k_resolve<<<cdiv(npix, 128), 128, 0, st>>>(p.paths.acc, npix, ns, accum, p0);
            { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "launch of k_resolve failed: %s", cudaGetErrorString(e_))); }
            launches++;
        }
    }
    if (adaptive) {
        // RenderPixel's second loop (main.cpp:246-258), jitter x 1.0. The reference takes one more sample per pixel and iteration; here every
        // still-active pixel gets its next K samples at once (K = 2, 4, 8, ...: at most twice what the pixel ends up using) and
        // k_adaptive_update replays the reference's per-sample decisions over them. Samples past a pixel's stopping point are discarded,
        // rays included (per-path ray counts), so colours, sample counts and ray_count are those of the one-at-a-time loop.
        uint32_t *lists[2][2] = {{sc->ad_u32 + pixel_count, sc->ad_u32 + 2 * (size_t)pixel_count},
                                 {sc->ad_u32 + 3 * (size_t)pixel_count, sc->ad_u32 + 4 * (size_t)pixel_count}};
        uint32_t *d_count = sc->ad_u32 + 5 * (size_t)pixel_count;
        unsigned long long *d_rays = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(d_count + 1) + 7u) & ~(uintptr_t)7u);
        CKR(cudaMemsetAsync(d_rays, 0, 8, st));
        k_adaptive_init<<<cdiv(pixel_count, 256), 256, 0, st>>>(pixel_count, d_ids, pixel_begin, lists[0][0], lists[0][1], d_nsamples, max_s);
        launches++;
        adaptive_first_pass_rays = sc->stats.closest_rays + sc->stats.shadow_rays;
        p.paths.ray_cnt = p.ray_cnt;
        uint32_t n_active = pixel_count;
        int cur = 0;
        uint32_t K = 2;
        for (uint32_t samp = params->min_samples; samp < max_s && n_active > 0;) {
            // chunk = the doubling schedule, or whatever it takes to put ~4 M samples in flight (a 720x480 frame cannot fill the chip with less)
            const uint32_t K_fill = (uint32_t)std::min<uint64_t>(((4ull << 20) + n_active - 1) / n_active, 1u << 20);
            const uint32_t Kc = std::max(1u, std::min(std::min(std::max(K, K_fill), max_s - samp), p.capacity / std::max(1u, std::min(n_active, p.capacity))));
            const uint32_t na_max = std::max(1u, p.capacity / Kc);
            CKR(cudaMemsetAsync(d_count, 0, 4, st));
            for (uint32_t a0 = 0; a0 < n_active; a0 += na_max) {   // more active samples than pool slots: chunks
                const uint32_t na = std::min(na_max, n_active - a0);
                PrimaryGen gen;
                memset(&gen, 0, sizeof(gen));
                gen.cam = dcam; gen.base_seed = prm.base_seed; gen.pixel_ids = lists[cur][0] + a0; gen.n_slots = na * Kc; gen.spp = Kc; gen.width = width;
                gen.pixel_begin = 0; gen.pixel_local0 = 0; gen.sample_begin = sample_begin + samp; gen.jitter_scale = 1.0f; gen.enabled = 1;
                rc = run_waves(sc, prm, na * Kc, flags, &launches, &gen);
                if (rc) { p.paths.ray_cnt = nullptr; return done(rc); }
                k_adaptive_update<<<cdiv(na, 128), 128, 0, st>>>(p.paths.acc, p.ray_cnt, na, samp, Kc, max_s, lists[cur][0] + a0, lists[cur][1] + a0, accum, sc->scratch, (size_t)pixel_count,
                                                                d_nsamples, lists[cur ^ 1][0], lists[cur ^ 1][1], d_count, d_rays);
                { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { p.paths.ray_cnt = nullptr; return done(fail(RT_ERR_CUDA, "launch of k_adaptive_update failed: %s", cudaGetErrorString(e_))); } }
                launches++;
            }
            CKR(cudaMemcpyAsync(&n_active, d_count, 4, cudaMemcpyDeviceToHost, st));
            CKR(cudaStreamSynchronize(st));
            cur ^= 1;
            samp += Kc;
            K = std::min(K * 2u, 1u << 20);
        }
        p.paths.ray_cnt = nullptr;
        CKR(cudaMemcpyAsync(&adaptive_accepted_rays, d_rays, 8, cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
        adaptive_counts = true;
        sc->last_adaptive_pixels = pixel_count;
    }
    k_finalize<<<cdiv(pixel_count, 128), 128, 0, st>>>(accum, pixel_count, sample_count, d_nsamples, flags & 3u, (float4 *)out_dev, d_ids, pixel_begin);
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "launch of k_finalize failed: %s", cudaGetErrorString(e_))); }
    launches++;
    CKR(cudaEventRecord(sc->ev1, st));
    TraceCounters tc;
    CKR(cudaMemcpyAsync(&tc, p.tcount, sizeof(tc), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0;
    CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms;
    sc->stats.kernel_launches = launches;
    { int rc_ = collect_wave_times(sc); if (rc_) return done(rc_); }
    if (out_counters) {
        out_counters->ray_count = adaptive_counts ? adaptive_first_pass_rays + adaptive_accepted_rays : sc->stats.closest_rays + sc->stats.shadow_rays;
        out_counters->sphere_check_count = tc.sphere_checks;
        out_counters->mesh_check_count = tc.cluster_checks;
    }
    if (user_stream) {      // make the caller's stream see the result
        CKR(cudaEventRecord(sc->ev1, st));
        CKR(cudaStreamWaitEvent(user_stream, sc->ev1, 0));
    }
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_render_device(rt_scene *scene, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                                const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                                uint32_t sample_count, uint32_t flags, float *out_rgba_device, void *stream, rt_counters *out_counters) {
    g_err.clear();
    return render_impl(scene, cam, params, width, height, pixel_ids, pixel_begin, pixel_count, sample_begin, sample_count, flags,
                       out_rgba_device, (cudaStream_t)stream, out_counters);
}

extern "C" int rt_render(rt_scene *scene, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                         const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                         uint32_t sample_count, uint32_t flags, float *out_rgba_host, rt_counters *out_counters) {
    g_err.clear();
    if (!scene || !out_rgba_host) return fail(RT_ERR_ARG, "null argument");
    if (flags & RT_OUT_FULLFRAME) return fail(RT_ERR_ARG, "RT_OUT_FULLFRAME is a device-output mode");
    CK(cudaSetDevice(scene->device));
    int rc = grow(scene, &scene->out_stage, &scene->out_stage_cap, (size_t)pixel_count * 4);
    if (rc) return rc;
    float *d_out = scene->out_stage;
    rc = render_impl(scene, cam, params, width, height, pixel_ids, pixel_begin, pixel_count, sample_begin, sample_count, flags,
                         d_out, nullptr, out_counters);
    if (rc == RT_OK && pixel_count) {
        cudaError_t e = cudaMemcpyAsync(out_rgba_host, d_out, (size_t)pixel_count * 16, cudaMemcpyDeviceToHost, scene->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(scene->stream);
        if (e != cudaSuccess) rc = fail(RT_ERR_CUDA, "framebuffer download failed: %s", cudaGetErrorString(e));
        scene->stats.d2h_bytes += (uint64_t)pixel_count * 16;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// rt_trace_rays / rt_trace_primary / rt_trace_color
// ---------------------------------------------------------------------------------------------
extern "C" int rt_trace_rays(rt_scene *sc, const rt_params *params, const rt_ray *rays, uint64_t n, int mode, rt_hit *out_hits,
                             rt_counters *out_counters) {
    g_err.clear();
    if (!sc || !params || (n && (!rays || !out_hits))) return fail(RT_ERR_ARG, "null argument");
    if (mode != RT_TRACE_CLOSEST && mode != RT_TRACE_ANY && mode != RT_TRACE_BRUTE) return fail(RT_ERR_ARG, "unknown trace mode %d", mode);
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    memset(&sc->stats, 0, sizeof(sc->stats));
    const uint32_t chunk = 1u << 20;
    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    float *d_rays; RayQueue q; HitRec *hits; ApiHit *api; TraceCounters *tc; float4 *rad, *acc; uint32_t *cnt;
    uint32_t cap = (uint32_t)std::min<uint64_t>(n ? n : 1, chunk);
    CKR(tmp.alloc(&d_rays, 6 * (size_t)cap)); CKR(tmp.alloc(&q.o, cap)); CKR(tmp.alloc(&q.d, cap)); CKR(tmp.alloc(&hits, cap));
    CKR(tmp.alloc(&api, cap)); CKR(tmp.alloc(&tc, 1)); CKR(tmp.alloc(&rad, cap)); CKR(tmp.alloc(&acc, cap)); CKR(tmp.alloc(&cnt, 4));
    CKR(cudaMemsetAsync(tc, 0, sizeof(TraceCounters), st));
    CKR(cudaEventRecord(sc->ev0, st));
    for (uint64_t b = 0; b < n; b += chunk) {
        uint32_t m = (uint32_t)std::min<uint64_t>(chunk, n - b);
        CKR(cudaMemcpyAsync(d_rays, rays + b, (size_t)m * sizeof(rt_ray), cudaMemcpyHostToDevice, st));
        WaveQueues w;
        memset(&w, 0, sizeof(w));
        w.next = cnt + 1; w.hits = hits; w.acc = acc; w.acc_extra = acc; w.rad = rad; w.shadow_o = q.o; w.shadow_dir = q.d; w.closest = q;
        w.n_shadow = cnt; w.shadow_stride = cap; set_fetch_knobs(w);
        int rc = RT_OK;
        if (mode == RT_TRACE_ANY) {
            k_rays_to_shadow_queue<<<cdiv(m, 256), 256, 0, st>>>(d_rays, m, q.o, q.d, rad, acc, cnt);
            w.closest_max = 0; w.n_lights = 1;
            rc = launch_trace_wave(sc, params->ray_bias, w, no_gen(), m, true, tc);
            k_occlusion_to_api<<<cdiv(m, 256), 256, 0, st>>>(acc, m, api);
        } else {
            k_upload_rays<<<cdiv(m, 256), 256, 0, st>>>(d_rays, m, q);
            w.closest_max = m; w.n_lights = 0;
            if (mode == RT_TRACE_BRUTE) k_trace_brute<<<cdiv(m, 128), 128, 0, st>>>(sc->d, params->ray_bias, q, m, hits);
            else rc = launch_trace_wave(sc, params->ray_bias, w, no_gen(), m, true, tc);
            k_hits_to_api<<<cdiv(m, 256), 256, 0, st>>>(sc->d, params->ray_bias, q, hits, m, api);
        }
        if (rc) return done(rc);
        { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "trace launch failed: %s", cudaGetErrorString(e_))); }
        CKR(cudaMemcpyAsync(out_hits + b, api, (size_t)m * sizeof(ApiHit), cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
        sc->stats.kernel_launches += 3;
        if (mode == RT_TRACE_ANY) sc->stats.shadow_rays += m; else sc->stats.closest_rays += m;
    }
    CKR(cudaEventRecord(sc->ev1, st));
    TraceCounters htc;
    CKR(cudaMemcpyAsync(&htc, tc, sizeof(htc), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0; CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms;
    if (out_counters) { out_counters->ray_count = n; out_counters->sphere_check_count = htc.sphere_checks; out_counters->mesh_check_count = htc.cluster_checks; }
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_trace_primary(rt_scene *sc, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                                const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                                uint32_t sample_count, rt_ray *out_rays, rt_hit *out_hits) {
    g_err.clear();
    if (!sc || !cam) return fail(RT_ERR_ARG, "null argument");
    int rc = check_params(params);
    if (rc) return rc;
    if (!width || !height) return fail(RT_ERR_ARG, "empty frame");
    const uint64_t frame = (uint64_t)width * height;
    if (!pixel_ids && (uint64_t)pixel_begin + pixel_count > frame) return fail(RT_ERR_ARG, "pixel range exceeds the frame");
    if (pixel_ids) for (uint32_t k = 0; k < pixel_count; ++k) if (pixel_ids[k] >= frame) return fail(RT_ERR_ARG, "pixel id %u out of range", pixel_ids[k]);
    if (!pixel_count || !sample_count) return RT_OK;
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    memset(&sc->stats, 0, sizeof(sc->stats));
    const uint32_t limit = 1u << 20;
    const uint32_t spp_chunk = std::min(sample_count, limit);
    const uint32_t pix_per_batch = std::max(1u, std::min(pixel_count, limit / spp_chunk));
    rc = ensure_pool(sc, pix_per_batch * spp_chunk, params->bounce_depth);
    if (rc) return rc;
    Pool &p = sc->pool;
    DevParams prm = to_dev_params(params);
    DevCamera dcam = to_dev_camera(cam);
    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    uint32_t *d_ids = nullptr; float *d_rays; ApiHit *api;
    const uint32_t cap = pix_per_batch * spp_chunk;
    CKR(tmp.alloc(&d_rays, 6 * (size_t)cap)); CKR(tmp.alloc(&api, cap));
    if (pixel_ids) { CKR(tmp.alloc(&d_ids, pixel_count)); CKR(cudaMemcpyAsync(d_ids, pixel_ids, (size_t)pixel_count * 4, cudaMemcpyHostToDevice, st)); }
    CKR(cudaEventRecord(sc->ev0, st));
    for (uint32_t p0 = 0; p0 < pixel_count; p0 += pix_per_batch) {
        const uint32_t npix = std::min(pix_per_batch, pixel_count - p0);
        for (uint32_t s0 = 0; s0 < sample_count; s0 += spp_chunk) {
            const uint32_t ns = std::min(spp_chunk, sample_count - s0);
            const uint32_t m = npix * ns;
            PrimaryGen gen;
            memset(&gen, 0, sizeof(gen));
            gen.cam = dcam; gen.base_seed = prm.base_seed; gen.pixel_ids = d_ids; gen.n_slots = m; gen.spp = ns; gen.width = width;
            gen.pixel_begin = pixel_begin; gen.pixel_local0 = p0; gen.sample_begin = sample_begin + s0; gen.jitter_scale = 0.5f; gen.enabled = 1;
            k_raygen<<<cdiv(m, 256), 256, 0, st>>>(gen, prm, p.paths, p.q[0], p.counts);
            if (out_hits) {
                WaveQueues w = wave_queues(sc, 0, m);
                w.n_closest = nullptr; w.n_lights = 0;
                rc = launch_trace_wave(sc, prm.ray_bias, w, no_gen(), m, false, p.tcount);
                if (rc) return done(rc);
                k_hits_to_api<<<cdiv(m, 256), 256, 0, st>>>(sc->d, prm.ray_bias, p.q[0], p.hits, m, api);
                sc->stats.kernel_launches += 2; sc->stats.closest_rays += m;
            }
            if (out_rays) k_queue_to_rays<<<cdiv(m, 256), 256, 0, st>>>(p.q[0], m, d_rays);
            { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "primary launch failed: %s", cudaGetErrorString(e_))); }
            sc->stats.kernel_launches += 1;
            // entries are pixel-major over the WHOLE call: (p0 + k) * sample_count + s0 + s
            for (uint32_t k = 0; k < npix; ++k) {
                size_t dst = (size_t)(p0 + k) * sample_count + s0, src = (size_t)k * ns;
                if (out_hits) CKR(cudaMemcpyAsync(out_hits + dst, api + src, (size_t)ns * sizeof(ApiHit), cudaMemcpyDeviceToHost, st));
                if (out_rays) CKR(cudaMemcpyAsync(out_rays + dst, d_rays + 6 * src, (size_t)ns * sizeof(rt_ray), cudaMemcpyDeviceToHost, st));
                if (ns == sample_count) {   // contiguous: one copy covers the whole batch
                    if (out_hits) CKR(cudaMemcpyAsync(out_hits + dst, api + src, (size_t)ns * npix * sizeof(ApiHit), cudaMemcpyDeviceToHost, st));
                    if (out_rays) CKR(cudaMemcpyAsync(out_rays + dst, d_rays + 6 * src, (size_t)ns * npix * sizeof(rt_ray), cudaMemcpyDeviceToHost, st));
                    break;
                }
            }
            CKR(cudaStreamSynchronize(st));
        }
    }
    CKR(cudaEventRecord(sc->ev1, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0; CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms;
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_trace_color(rt_scene *sc, const rt_params *params, const rt_ray *rays, const uint64_t *seeds, uint64_t n,
                              float *out_rgba, rt_counters *out_counters) {
    g_err.clear();
    if (!sc || (n && (!rays || !seeds || !out_rgba))) return fail(RT_ERR_ARG, "null argument");
    int rc = check_params(params);
    if (rc) return rc;
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    memset(&sc->stats, 0, sizeof(sc->stats));
    if (out_counters) memset(out_counters, 0, sizeof(*out_counters));
    if (!n) return RT_OK;
    rc = ensure_spec_table(sc, params->spec_samples);
    if (rc) return rc;
    const uint32_t chunk = (uint32_t)std::min<uint64_t>(n, 1u << 20);
    rc = ensure_pool(sc, chunk, params->bounce_depth);
    if (rc) return rc;
    Pool &p = sc->pool;
    DevParams prm = to_dev_params(params);
    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    float *d_rays; uint64_t *d_seeds;
    CKR(tmp.alloc(&d_rays, 6 * (size_t)chunk)); CKR(tmp.alloc(&d_seeds, chunk));
    CKR(cudaMemsetAsync(p.tcount, 0, sizeof(TraceCounters), st));
    uint64_t launches = 0;
    CKR(cudaEventRecord(sc->ev0, st));
    for (uint64_t b = 0; b < n; b += chunk) {
        uint32_t m = (uint32_t)std::min<uint64_t>(chunk, n - b);
        CKR(cudaMemcpyAsync(d_rays, rays + b, (size_t)m * sizeof(rt_ray), cudaMemcpyHostToDevice, st));
        CKR(cudaMemcpyAsync(d_seeds, seeds + b, (size_t)m * 8, cudaMemcpyHostToDevice, st));
        k_paths_from_rays<<<cdiv(m, 256), 256, 0, st>>>(prm, p.paths, p.q[0], m, d_rays, d_seeds, p.counts);
        { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "launch of k_paths_from_rays failed: %s", cudaGetErrorString(e_))); }
        launches++;
        rc = run_waves(sc, prm, m, RT_FLAG_COUNTERS, &launches);
        if (rc) return done(rc);
        CKR(cudaMemcpyAsync(out_rgba + 4 * b, p.paths.acc, (size_t)m * 16, cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
    }
    CKR(cudaEventRecord(sc->ev1, st));
    TraceCounters tc;
    CKR(cudaMemcpyAsync(&tc, p.tcount, sizeof(tc), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0; CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms; sc->stats.kernel_launches = launches;
    if (out_counters) {
        out_counters->ray_count = sc->stats.closest_rays + sc->stats.shadow_rays;
        out_counters->sphere_check_count = tc.sphere_checks; out_counters->mesh_check_count = tc.cluster_checks;
    }
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_get_stats(const rt_scene *sc, rt_stats *out) {
    if (!sc || !out) return fail(RT_ERR_ARG, "null argument");
    *out = sc->stats;
    return RT_OK;
}

extern "C" int rt_get_hierarchy_info(const rt_scene *sc, uint64_t out[8]) {
    if (!sc || !out) return fail(RT_ERR_ARG, "null argument");
    memcpy(out, sc->info, sizeof(sc->info));
    return RT_OK;
}

extern "C" int rt_rng_kat(int device, uint64_t seed, uint32_t n, uint64_t *out_host) {
    g_err.clear();
    if (!out_host) return fail(RT_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    uint64_t *d;
    CK(cudaMalloc((void **)&d, std::max(1u, n) * 8));
    k_rng_kat<<<1, 1>>>(seed, n, d);
    cudaError_t e = cudaMemcpy(out_host, d, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, "rng kat failed: %s", cudaGetErrorString(e));
    return RT_OK;
}

// Per-pixel sample counts (RenderPixel's final `samp`, main.cpp:262) of the last RT_FLAG_ADAPTIVE render on this scene.
extern "C" int rt_get_sample_counts(rt_scene *sc, uint32_t *out_host, uint32_t n) {
    g_err.clear();
    if (!sc || !out_host) return fail(RT_ERR_ARG, "null argument");
    if (n > sc->last_adaptive_pixels || !sc->ad_u32) return fail(RT_ERR_STATE, "no adaptive render of >= %u pixels on this scene", n);
    CK(cudaSetDevice(sc->device));
    CK(cudaMemcpyAsync(out_host, sc->ad_u32, (size_t)n * 4, cudaMemcpyDeviceToHost, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// rt_tonemap_device / rt_tonemap: WriteFramebufferImage's tone map + Color_Pack (main.cpp:101-127) without the PNG
// ---------------------------------------------------------------------------------------------
extern "C" int rt_tonemap_device(int device, const float *rgba_device, uint32_t width, uint32_t height, uint8_t *out_rgba8_device,
                                 float *out_scene_luma_host, void *stream) {
    g_err.clear();
    if (!rgba_device || !out_rgba8_device) return fail(RT_ERR_ARG, "null argument");
    const uint64_t n64 = (uint64_t)width * height;
    if (n64 == 0 || n64 > 0xFFFFFFFFull) return fail(RT_ERR_ARG, "bad frame size");
    const uint32_t n = (uint32_t)n64;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t nb = std::min(cdiv(n, 256), 1024u);
    double *partial; float *luma;
    CK(cudaMalloc((void **)&partial, nb * sizeof(double) + sizeof(float)));
    luma = (float *)(partial + nb);
    k_luma_partial<<<nb, 256, 0, st>>>((const float4 *)rgba_device, n, partial);
    k_luma_final<<<1, 32, 0, st>>>(partial, nb, n, luma);
    k_tonemap_pack<<<cdiv(n, 256), 256, 0, st>>>((const float4 *)rgba_device, n, luma, (uchar4 *)out_rgba8_device);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && out_scene_luma_host) e = cudaMemcpyAsync(out_scene_luma_host, luma, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(partial);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, "tone map failed: %s", cudaGetErrorString(e));
    return RT_OK;
}

extern "C" int rt_tonemap(int device, const float *rgba_host, uint32_t width, uint32_t height, uint8_t *out_rgba8_host, float *out_scene_luma) {
    g_err.clear();
    if (!rgba_host || !out_rgba8_host) return fail(RT_ERR_ARG, "null argument");
    const size_t n = (size_t)width * height;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    float *d_in; uint8_t *d_out;
    CK(cudaMalloc((void **)&d_in, std::max<size_t>(1, n) * 16));
    if (cudaMalloc((void **)&d_out, std::max<size_t>(1, n) * 4) != cudaSuccess) { cudaFree(d_in); return fail(RT_ERR_NOMEM, "out of device memory"); }
    int rc = RT_OK;
    if (cudaMemcpy(d_in, rgba_host, n * 16, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(RT_ERR_CUDA, "upload failed");
    if (rc == RT_OK) rc = rt_tonemap_device(device, d_in, width, height, d_out, out_scene_luma, nullptr);
    if (rc == RT_OK && cudaMemcpy(out_rgba8_host, d_out, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(RT_ERR_CUDA, "download failed");
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// rt_build_group_hierarchy: BuildHierarchy (bsphere.cpp:379-444) on the GPU, bit-identical output
// ---------------------------------------------------------------------------------------------
extern "C" int rt_build_group_hierarchy(int device, const float *positions, uint32_t n_positions, uint32_t n_groups, const uint32_t *group_first,
                                        const uint32_t *idx_positions, rt_bsphere *out_spheres, int32_t *out_sphere_group, uint32_t *out_count) {
    g_err.clear();
    if (out_count) *out_count = 0;
    if (n_groups == 0) return RT_OK;
    if (!positions || !group_first || !idx_positions || !out_spheres || !out_sphere_group) return fail(RT_ERR_ARG, "null argument");
    if (n_groups > 65535u) return fail(RT_ERR_ARG, "at most 65535 mesh groups (got %u)", n_groups);
    const uint64_t n_idx = group_first[n_groups];
    for (uint32_t g = 0; g < n_groups; ++g) if (group_first[g + 1] <= group_first[g]) return fail(RT_ERR_ARG, "group %u is empty", g);
    for (uint64_t i = 0; i < n_idx; ++i) if (idx_positions[i] >= n_positions) return fail(RT_ERR_ARG, "vertex index out of range at %llu", (unsigned long long)i);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    const uint32_t total = 2 * n_groups - 1;
    DevArena mem;
    auto done = [&](int r) { mem.release(); return r; };
#define CKG(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    float *d_pos, *d_pts; uint32_t *d_gf, *d_idx, *d_list[2]; GSphere *d_S; int32_t *d_c0, *d_c1; unsigned long long *d_best;
    CKG(mem.alloc(&d_pos, 3 * (size_t)n_positions)); CKG(mem.alloc(&d_pts, 3 * (size_t)n_idx)); CKG(mem.alloc(&d_gf, (size_t)n_groups + 1));
    CKG(mem.alloc(&d_idx, (size_t)n_idx)); CKG(mem.alloc(&d_list[0], n_groups)); CKG(mem.alloc(&d_list[1], n_groups));
    CKG(mem.alloc(&d_S, total)); CKG(mem.alloc(&d_c0, total)); CKG(mem.alloc(&d_c1, total)); CKG(mem.alloc(&d_best, 1));
    CKG(cudaMemcpy(d_pos, positions, 12 * (size_t)n_positions, cudaMemcpyHostToDevice));
    CKG(cudaMemcpy(d_gf, group_first, 4 * ((size_t)n_groups + 1), cudaMemcpyHostToDevice));
    CKG(cudaMemcpy(d_idx, idx_positions, 4 * (size_t)n_idx, cudaMemcpyHostToDevice));
    CKG(cudaMemset(d_c0, 0xFF, 4 * (size_t)total)); CKG(cudaMemset(d_c1, 0xFF, 4 * (size_t)total));
    CKG(cudaMemset(d_best, 0xFF, 8));
    {
        std::vector<uint32_t> iota(n_groups);
        for (uint32_t g = 0; g < n_groups; ++g) iota[g] = g;
        CKG(cudaMemcpy(d_list[0], iota.data(), 4 * (size_t)n_groups, cudaMemcpyHostToDevice));
    }
    k_group_leaf_spheres<<<cdiv(n_groups, 32), 32>>>(d_pos, d_gf, d_idx, n_groups, d_pts, d_S);
    int cur = 0;
    uint32_t created = n_groups;
    for (uint32_t m = n_groups; m >= 2; --m) {          // every merge removes two spheres and appends one
        k_group_pair_min<<<m, 256>>>(m, d_list[cur], d_S, d_best);
        k_group_apply_merge<<<1, 1024>>>(m, d_list[cur], d_list[cur ^ 1], d_S, d_c0, d_c1, created, d_best);
        created++;
        cur ^= 1;
    }
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "group hierarchy launch failed: %s", cudaGetErrorString(e_))); }
    std::vector<GSphere> S(total); std::vector<int32_t> c0(total), c1(total);
    uint32_t root = 0;
    CKG(cudaMemcpy(S.data(), d_S, sizeof(GSphere) * (size_t)total, cudaMemcpyDeviceToHost));
    CKG(cudaMemcpy(c0.data(), d_c0, 4 * (size_t)total, cudaMemcpyDeviceToHost));
    CKG(cudaMemcpy(c1.data(), d_c1, 4 * (size_t)total, cudaMemcpyDeviceToHost));
    CKG(cudaMemcpy(&root, d_list[cur], 4, cudaMemcpyDeviceToHost));
    // FlattenHierarchyTree (bsphere.cpp:328-350): pre-order, child index 0 == leaf sentinel -- pure index bookkeeping, done on the host
    std::vector<uint32_t> stack, order, slot_of(total, 0);
    stack.push_back(root);
    while (!stack.empty()) {
        uint32_t n = stack.back(); stack.pop_back();
        if (n >= total) return done(fail(RT_ERR_STATE, "group hierarchy corrupt"));
        slot_of[n] = (uint32_t)order.size(); order.push_back(n);
        if (c0[n] >= 0) { stack.push_back((uint32_t)c1[n]); stack.push_back((uint32_t)c0[n]); }
        if (order.size() > total) return done(fail(RT_ERR_STATE, "group hierarchy corrupt"));
    }
    for (size_t k = 0; k < order.size(); ++k) {
        uint32_t n = order[k];
        out_spheres[k].center[0] = S[n].x; out_spheres[k].center[1] = S[n].y; out_spheres[k].center[2] = S[n].z; out_spheres[k].radius = S[n].r;
        out_spheres[k].c0 = c0[n] >= 0 ? slot_of[c0[n]] : 0;
        out_spheres[k].c1 = c1[n] >= 0 ? slot_of[c1[n]] : 0;
        out_sphere_group[k] = c0[n] >= 0 ? -1 : (int32_t)n;
    }
    if (out_count) *out_count = (uint32_t)order.size();
    return done(RT_OK);
#undef CKG
}

// ---------------------------------------------------------------------------------------------
// rt_calculate_tangents / rt_height_to_normal_map: the reference's load-time preprocessing (mesh.h:59-129, texture.cpp:85-144)
// ---------------------------------------------------------------------------------------------
extern "C" int rt_calculate_tangents(int device, const float *positions, uint32_t n_positions, const float *texcoords, uint32_t n_texcoords,
                                     uint32_t n_normals, uint32_t n_groups, const uint32_t *group_first, const uint32_t *idx_positions,
                                     const uint32_t *idx_texcoords, const uint32_t *idx_normals, const uint8_t *group_has_bump, float *out_tangents) {
    g_err.clear();
    if (!out_tangents) return fail(RT_ERR_ARG, "null argument");
    memset(out_tangents, 0, sizeof(float) * 3 * (size_t)n_normals);
    if (n_groups == 0 || n_normals == 0) return RT_OK;
    if (!positions || !texcoords || !group_first || !idx_positions || !idx_texcoords || !idx_normals || !group_has_bump) return fail(RT_ERR_ARG, "null argument");
    const uint64_t n_idx = group_first[n_groups];
    if (n_idx % 3 || n_idx / 3 > 200000000ull) return fail(RT_ERR_ARG, "bad index count");
    for (uint64_t i = 0; i < n_idx; ++i)
        if (idx_positions[i] >= n_positions || idx_texcoords[i] >= n_texcoords || idx_normals[i] >= n_normals) return fail(RT_ERR_ARG, "vertex index out of range at %llu", (unsigned long long)i);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    const uint32_t n_tris = (uint32_t)(n_idx / 3);
    uint32_t n_pad = BITONIC_TILE;
    while (n_pad < n_idx) n_pad <<= 1;
    DevArena mem;
    auto done = [&](int r) { mem.release(); return r; };
#define CKT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    float *d_pos, *d_tc, *d_tan; uint32_t *d_ip, *d_it, *d_in, *d_gf, *vals; uint8_t *d_hb; float4 *tri_tan; uint64_t *keys;
    CKT(mem.alloc(&d_pos, 3 * (size_t)n_positions)); CKT(mem.alloc(&d_tc, 2 * (size_t)n_texcoords)); CKT(mem.alloc(&d_tan, 3 * (size_t)n_normals));
    CKT(mem.alloc(&d_ip, (size_t)n_idx)); CKT(mem.alloc(&d_it, (size_t)n_idx)); CKT(mem.alloc(&d_in, (size_t)n_idx)); CKT(mem.alloc(&d_gf, (size_t)n_groups + 1));
    CKT(mem.alloc(&d_hb, n_groups)); CKT(mem.alloc(&tri_tan, n_tris)); CKT(mem.alloc(&keys, n_pad)); CKT(mem.alloc(&vals, n_pad));
    CKT(cudaMemcpy(d_pos, positions, 12 * (size_t)n_positions, cudaMemcpyHostToDevice));
    CKT(cudaMemcpy(d_tc, texcoords, 8 * (size_t)n_texcoords, cudaMemcpyHostToDevice));
    CKT(cudaMemcpy(d_ip, idx_positions, 4 * (size_t)n_idx, cudaMemcpyHostToDevice));
    CKT(cudaMemcpy(d_it, idx_texcoords, 4 * (size_t)n_idx, cudaMemcpyHostToDevice));
    CKT(cudaMemcpy(d_in, idx_normals, 4 * (size_t)n_idx, cudaMemcpyHostToDevice));
    CKT(cudaMemcpy(d_gf, group_first, 4 * ((size_t)n_groups + 1), cudaMemcpyHostToDevice));
    CKT(cudaMemcpy(d_hb, group_has_bump, n_groups, cudaMemcpyHostToDevice));
    CKT(cudaMemset(d_tan, 0, 12 * (size_t)n_normals));
    TangentInput in; in.positions = d_pos; in.texcoords = d_tc; in.idx_p = d_ip; in.idx_t = d_it; in.idx_n = d_in; in.group_first = d_gf;
    in.group_has_bump = d_hb; in.n_groups = n_groups; in.n_tris = n_tris;
    k_tri_tangents<<<cdiv((n_pad + 2) / 3 + 1, 256), 256>>>(in, tri_tan, keys, vals, n_pad);
    k_bitonic_shared<<<n_pad / BITONIC_TILE, 1024>>>(keys, vals, 2, BITONIC_TILE, 0);
    for (uint64_t k = 2ull * BITONIC_TILE; k <= n_pad; k <<= 1) {
        for (uint32_t j = (uint32_t)(k >> 1); j >= BITONIC_TILE; j >>= 1) k_bitonic_global<<<cdiv(n_pad, 256), 256>>>(keys, vals, n_pad, j, (uint32_t)k);
        k_bitonic_shared<<<n_pad / BITONIC_TILE, 1024>>>(keys, vals, (uint32_t)k, (uint32_t)k, 1);
    }
    k_sum_tangents<<<cdiv(n_pad, 256), 256>>>(keys, vals, n_pad, tri_tan, d_tan);
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "tangent launch failed: %s", cudaGetErrorString(e_))); }
    CKT(cudaMemcpy(out_tangents, d_tan, 12 * (size_t)n_normals, cudaMemcpyDeviceToHost));
    return done(RT_OK);
#undef CKT
}

extern "C" int rt_height_to_normal_map(int device, uint32_t size_x, uint32_t size_y, const uint8_t *height_host, uint8_t *out_rgb_host) {
    g_err.clear();
    if (!height_host || !out_rgb_host || !size_x || !size_y) return fail(RT_ERR_ARG, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    const size_t n = (size_t)size_x * size_y;
    float lut[256]; host_srgb_lut(lut);
    uint8_t *d_h = nullptr, *d_o = nullptr; float *d_lut = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_h, n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_o, 3 * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_lut, sizeof(lut));
    if (e == cudaSuccess) e = cudaMemcpy(d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_h, height_host, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        dim3 grid(cdiv(size_x, 128), size_y);
        k_height_to_normal<<<grid, 128>>>(size_x, size_y, d_h, d_lut, d_o);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_rgb_host, d_o, 3 * n, cudaMemcpyDeviceToHost);
    if (d_h) cudaFree(d_h);
    if (d_o) cudaFree(d_o);
    if (d_lut) cudaFree(d_lut);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, "height map conversion failed: %s", cudaGetErrorString(e));
    return RT_OK;
}
