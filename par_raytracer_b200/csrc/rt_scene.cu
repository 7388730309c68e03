// rt_scene.cu -- errors, scene upload, GPU hierarchy build (kernels in rt_build.cuh), introspection.
// Host code here only moves data, sizes launches and builds small decode tables. There is no CPU fallback.
#include "rt_internal.h"
#include "rt_build.cuh"

thread_local std::string g_rt_err;

int rt_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_rt_err = buf;
    return code;
}

extern "C" const char *rt_last_error(void) { return g_err.c_str(); }
extern "C" int rt_abi_version(void) { return RT_ABI_VERSION; }
extern "C" int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

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

// ---------------------------------------------------------------------------------------------
// quantisation grid of QNode (host): 32764 steps across [lo, hi] per axis, at least one step of margin below lo.
// The kernel decodes plane q as mid + (32768 + q) * step from these two FLOATS, so the grid is built from exactly them.
// A scene far from the origin relative to its size rounds `mid` coarsely: the step is widened until the grid covers [lo, hi].
// Returns false when no grid can (non-finite bounds).
// ---------------------------------------------------------------------------------------------
bool rt_place_quant_grid(double lo, double hi, float *step_out, float *mid_out, double *base_out) {
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
    for (int a = 0; a < 3; ++a) { double b; if (!rt_place_quant_grid(lo[a], hi[a], &step[a], &mid[a], &b)) *ok = 0; }
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
    TriRec *tris; uint32_t *tri_rank, *tri_vertex0; int32_t *tri_object; float4 *tri_uv, *tri_nrm, *tri_tan = nullptr; uint8_t *tri_mat;
    CKB(sc->mem.alloc(&tris, n)); CKB(sc->mem.alloc(&tri_rank, n)); CKB(sc->mem.alloc(&tri_vertex0, n));
    CKB(sc->mem.alloc(&tri_object, n)); CKB(sc->mem.alloc(&tri_mat, n)); CKB(sc->mem.alloc(&tri_uv, 2 * (size_t)n)); CKB(sc->mem.alloc(&tri_nrm, 3 * (size_t)n));
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
            if (!rt_place_quant_grid(lo3[a], hi3[a], &sc->d.qstep[a], &sc->d.qmid[a], &qb[a])) grid_ok = false;     // non-finite extents
            qs[a] = (double)sc->d.qstep[a];
        }
    }
    if (!grid_ok && (sc->bounds == RT_BOUNDS_QBOX || sc->bounds == RT_BOUNDS_QBOX4)) sc->bounds = RT_BOUNDS_BOX;   // float boxes need no grid (inf / NaN vertices behave as they do there)
    // only the node array of the selected child bound is built (RT_B200_BOUNDS)
    HNode *nodes = nullptr; BNode *bnodes = nullptr; QNode *qnodes = nullptr; Q4Node *q4nodes = nullptr;
    uint32_t n_wide = 0, wide_levels = 0;
    if (sc->bounds == RT_BOUNDS_QBOX4 && kept_nodes > 0) {
        // collapse to four children per node, level by level from the root (kernels + comment in rt_build.cuh)
        WideTree w;
        CKB(tmp.alloc(&w.src, kept_nodes)); CKB(tmp.alloc(&w.child, 4 * (size_t)kept_nodes)); CKB(tmp.alloc(&w.ref, 4 * (size_t)kept_nodes));
        const int32_t root_id = (int32_t)(n_total - 1);
        CKB(cudaMemcpyAsync(w.src, &root_id, 4, cudaMemcpyHostToDevice, st));
        uint32_t first = 0, m = 1;
        while (m > 0) {
            const uint32_t nb = cdiv(m, SCAN_TILE);
            k_wide_expand<<<cdiv(m, 128), 128, 0, st>>>(m, first, t, w, flags); CKLB("k_wide_expand");
            k_scan_reduce<<<nb, SCAN_BLOCK, 0, st>>>(flags, m, bsums); CKLB("k_scan_reduce");
            k_scan_blocksums<<<1, 1024, 0, st>>>(bsums, nb, total); CKLB("k_scan_blocksums");
            k_scan_apply<<<nb, SCAN_BLOCK, 0, st>>>(flags, m, bsums, scan); CKLB("k_scan_apply");
            k_wide_assign<<<cdiv(m, 128), 128, 0, st>>>(m, first, first + m, t, w, scan, tri_offset); CKLB("k_wide_assign");
            uint64_t h_total = 0;
            CKB(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, st));
            CKB(cudaStreamSynchronize(st));
            first += m; m = (uint32_t)h_total; wide_levels++;
            if ((uint64_t)first + m > kept_nodes) return done(fail(RT_ERR_STATE, "4-wide collapse produced more nodes than the binary tree has"));
        }
        n_wide = first;
        if (3 * wide_levels + 2 > RT_STACK4_MAX) { sc->bounds = RT_BOUNDS_QBOX; n_wide = 0; }     // too deep for the 4-wide traversal stack: binary form
        else {
            CKB(sc->mem.alloc(&q4nodes, std::max(1u, n_wide)));
            k_wide_emit<<<cdiv(n_wide, 128), 128, 0, st>>>(n_wide, t, w, q4nodes, qb[0], qb[1], qb[2], qs[0], qs[1], qs[2]); CKLB("k_wide_emit");
        }
    }
    if (sc->bounds == RT_BOUNDS_QBOX4 && kept_nodes == 0) sc->bounds = RT_BOUNDS_QBOX;           // a scene of one cluster has no nodes at all
    if (sc->bounds == RT_BOUNDS_QBOX4) { /* node array emitted above */ }
    else if (sc->bounds == RT_BOUNDS_SPHERE) CKB(sc->mem.alloc(&nodes, std::max(1u, kept_nodes)));
    else if (sc->bounds == RT_BOUNDS_BOX) CKB(sc->mem.alloc(&bnodes, std::max(1u, kept_nodes)));
    else CKB(sc->mem.alloc(&qnodes, std::max(1u, kept_nodes)));
    if (n > 1) {
        k_slot_to_tri<<<cdiv(n, 256), 256, 0, st>>>(n, vals, tri_offset, slot_tri); CKLB("k_slot_to_tri");
        if (sc->bounds == RT_BOUNDS_SPHERE) { k_refit<<<cdiv((uint64_t)(n - 1) * 32, 256), 256, 0, st>>>(n, n_total, t, tri_offset, slot_tri, bin); CKLB("k_refit"); }
    }
    if (kept_nodes > 0 && sc->bounds == RT_BOUNDS_QBOX4) sc->d.root = 0;
    else if (kept_nodes > 0) {
        k_emit_nodes<<<cdiv(n - 1, 256), 256, 0, st>>>(n, n_total, t, tri_offset, kept_index, nodes, bnodes, qnodes, qb[0], qb[1], qb[2], qs[0], qs[1], qs[2]); CKLB("k_emit_nodes");
        sc->d.root = 0;
    } else {
        sc->d.root = -(int)(1u + 0u * 8u + n);     // the whole scene is one cluster (n <= RT_LEAF_MAX)
    }
    k_gather<<<cdiv(n, 256), 256, 0, st>>>(gin, vals, tri_offset, tris, tri_rank, tri_uv, tri_nrm, tri_tan, tri_vertex0, tri_object, tri_mat);
    CKLB("k_gather");
    CKB(cudaEventRecord(e1, st));
    CKB(cudaStreamSynchronize(st));
    float ms = 0;
    CKB(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);

    sc->d.nodes = nodes; sc->d.bnodes = bnodes; sc->d.qnodes = qnodes; sc->d.q4nodes = q4nodes; sc->d.tris = tris; sc->d.tri_rank = tri_rank; sc->d.tri_uv = tri_uv; sc->d.tri_nrm = tri_nrm;
    sc->d.tri_tan = tri_tan; sc->d.tri_vertex0 = tri_vertex0; sc->d.tri_object = tri_object; sc->d.tri_mat = tri_mat;
    sc->d.n_tris = n; sc->d.n_nodes = kept_nodes;
    {   // every sphere lies inside the root sphere: |c|_1 + r <= |c_root|_1 + sqrt(3) * 2 r_root + r_root
        float4 rs;
        CKB(cudaMemcpy(&rs, t.sphere + (n_total - 1), sizeof(rs), cudaMemcpyDeviceToHost));
        sc->d.cull_bound = fabsf(rs.x) + fabsf(rs.y) + fabsf(rs.z) + 4.5f * rs.w;
    }
    sc->info[0] = n; sc->info[1] = 0; sc->info[2] = kept_nodes; sc->info[3] = depth;
    sc->info[4] = sc->bounds == RT_BOUNDS_QBOX4 ? (uint64_t)n_wide * sizeof(Q4Node)
                : (uint64_t)kept_nodes * (sc->bounds == RT_BOUNDS_QBOX ? sizeof(QNode) : sc->bounds == RT_BOUNDS_BOX ? sizeof(BNode) : sizeof(HNode));
    if (sc->bounds == RT_BOUNDS_QBOX4) { sc->info[1] = n_wide; sc->info[3] = wide_levels; } sc->info[5] = (uint64_t)n * sizeof(TriRec);
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
    sc->pinned_out.release();
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
    {   // child bound of the GPU hierarchy; the build may fall back (non-finite extents -> float boxes, over-deep tree -> binary)
        const char *be = getenv("RT_B200_BOUNDS");
        sc->bounds = (be && strcmp(be, "sphere") == 0) ? RT_BOUNDS_SPHERE : (be && strcmp(be, "box") == 0) ? RT_BOUNDS_BOX
                   : (be && strcmp(be, "qbox4") == 0) ? RT_BOUNDS_QBOX4 : RT_BOUNDS_QBOX;
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
    { int rc_ = rt_render_configure(sc); if (rc_) return bail(rc_); }       // persistent-grid sizes for the bound the build settled on
    *out_scene = sc;
    return RT_OK;
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
