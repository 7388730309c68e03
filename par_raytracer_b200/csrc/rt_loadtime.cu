// rt_loadtime.cu -- the steps either side of the hot path: tone map + 8-bit pack, the reference's own BuildHierarchy over
// mesh groups, CalculateTangents, ConvertHeightMapToNormalMap -- each on the device, each pinned to the reference's output.
#include "rt_internal.h"
#include "rt_sort.cuh"
#include "rt_groups.cuh"
#include "rt_preprocess.cuh"
#include "rt_tonemap.cuh"

static void host_srgb_lut(float *lut) {                        // color.h:13-21 over texture.cpp:44-48's 256 inputs
    const float one_over_255 = 1.0f / 255.0f;
    for (int i = 0; i < 256; ++i) {
        volatile float srgb = (float)i * one_over_255;
        lut[i] = srgb <= 0.04045f ? srgb / 12.92f : powf((srgb + 0.055f) / 1.055f, 2.4f);
    }
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
