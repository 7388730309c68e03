// rt_preprocess.cuh -- SURVEY.md 8(f) #4: the reference's load-time preprocessing on the GPU.
//   CalculateTangents (mesh.h:59-129): per-triangle UV-delta tangents ACCUMULATED on the normal index in triangle order, then
//     normalised. Floating-point accumulation order is part of the result, so contributions are sorted by (normal index, corner
//     order) and each normal's run is summed sequentially -- the same additions in the same order as the reference's loop.
//   ConvertHeightMapToNormalMap + WriteNormal (texture.cpp:85-144): per texel; the sRGB encode's powf is evaluated in double and
//     rounded to float (glibc's float powf is correctly rounded in all but rare cases), then truncated like (u8)(x * 255).
#pragma once
#include "rt_common.cuh"
#include "rt_sort.cuh"       // bitonic sort kernels

struct TangentInput {
    const float *positions, *texcoords;
    const uint32_t *idx_p, *idx_t, *idx_n, *group_first;
    const uint8_t *group_has_bump;
    uint32_t n_groups, n_tris;
};

// one thread per triangle (global order = the reference's processing order): tangent + one sort key per corner
__global__ void k_tri_tangents(TangentInput in, float4 *tri_tan, uint64_t *keys, uint32_t *vals, uint32_t n_pad) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (3ull * t >= n_pad) return;
    bool valid = false;
    f3 tg = mk3(0, 0, 0);
    if (t < in.n_tris) {
        uint32_t g = find_group(in.group_first, in.n_groups, 3u * t);
        if (in.group_has_bump[g]) {                                                    // mesh.h:70-74
            f3 p0 = ld3(in.positions, in.idx_p[3 * (size_t)t]), p1 = ld3(in.positions, in.idx_p[3 * (size_t)t + 1]), p2 = ld3(in.positions, in.idx_p[3 * (size_t)t + 2]);
            const float *uv0 = in.texcoords + 2 * (size_t)in.idx_t[3 * (size_t)t], *uv1 = in.texcoords + 2 * (size_t)in.idx_t[3 * (size_t)t + 1],
                        *uv2 = in.texcoords + 2 * (size_t)in.idx_t[3 * (size_t)t + 2];
            f3 dp0 = p1 - p0, dp1 = p2 - p0;
            float d0x = uv1[0] - uv0[0], d0y = uv1[1] - uv0[1], d1x = uv2[0] - uv0[0], d1y = uv2[1] - uv0[1];
            float f = (d0x * d1y - d1x * d0y);
            if (!((double)f <= 1e-7)) {                                                // mesh.h:96: float against a double literal
                f = 1.0f / f;
                tg.x = f * (d1y * dp0.x - d0y * dp1.x);
                tg.y = f * (d1y * dp0.y - d0y * dp1.y);
                tg.z = f * (d1y * dp0.z - d0y * dp1.z);
                valid = true;
            }
        }
        tri_tan[t] = make_float4(tg.x, tg.y, tg.z, 0.0f);
    }
    for (uint32_t k = 0; k < 3; ++k) {
        uint64_t i = 3ull * t + k;
        if (i >= n_pad) break;
        if (valid) { keys[i] = ((uint64_t)in.idx_n[3 * (size_t)t + k] << 32) | (uint32_t)i; vals[i] = t; }
        else { keys[i] = ~0ull; vals[i] = ~0u; }
    }
}

// one thread per sorted entry; the first entry of each normal's run sums the run in order and normalises (mesh.h:115-128)
__global__ void k_sum_tangents(const uint64_t *keys, const uint32_t *vals, uint32_t n_pad, const float4 *tri_tan, float *tangents) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    uint64_t k = keys[i];
    if (k == ~0ull) return;
    uint32_t ni = (uint32_t)(k >> 32);
    if (i > 0 && (uint32_t)(keys[i - 1] >> 32) == ni) return;
    f3 acc = mk3(0, 0, 0);
    for (uint32_t j = i; j < n_pad && keys[j] != ~0ull && (uint32_t)(keys[j] >> 32) == ni; ++j) acc = acc + mk3(tri_tan[vals[j]]);
    acc = normalize3(acc);
    tangents[3 * (size_t)ni] = acc.x; tangents[3 * (size_t)ni + 1] = acc.y; tangents[3 * (size_t)ni + 2] = acc.z;
}

RT_DEVICE float linear_to_srgb_dev(float linear) {                                     // color.h:3-11
    if (linear <= 0.0031308f) return 12.92f * linear;
    return 1.055f * (float)pow((double)linear, (double)(1.0f / 2.4f)) - 0.055f;
}

__global__ void k_height_to_normal(uint32_t sx, uint32_t sy, const uint8_t *height, const float *srgb_lut, uint8_t *out_rgb) {   // texture.cpp:102-125
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= sx || y >= sy) return;
    uint32_t x1 = (x + 1) % sx, y1 = (y + 1) % sy;
    float h00 = srgb_lut[height[y * sx + x]], h10 = srgb_lut[height[y * sx + x1]], h01 = srgb_lut[height[y1 * sx + x]];
    float a = 2.5f;
    f3 n = normalize3(mk3((h01 - h00) * a, (h10 - h00) * a, 1.0f));
    n = (n + mk3(1.0f, 1.0f, 1.0f)) * 0.5f;                                            // texture.cpp:92-93
    uint8_t *o = out_rgb + 3 * ((size_t)y * sx + x);
    o[0] = (uint8_t)(linear_to_srgb_dev(n.x) * 255.0f);
    o[1] = (uint8_t)(linear_to_srgb_dev(n.y) * 255.0f);
    o[2] = (uint8_t)(linear_to_srgb_dev(n.z) * 255.0f);
}
