// rt_tonemap.cuh -- LogAverageLuma / WriteFramebufferImage's tone map / Color_Pack (main.cpp:78-127, color.h:94-111) on the device:
// the step right after Render(). Included by rt_loadtime.cu only.
#pragma once
#include "rt_common.cuh"

// ---- tone map + 8-bit pack: LogAverageLuma / WriteFramebufferImage / Color_Pack (main.cpp:78-127, color.h:94-111) --------
RT_DEVICE float color_luma(float4 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }     // color.h:94-97

// sum of logf(0.01 + luma) over pixels with luma > 0 (main.cpp:84-97). The reference adds 32-bit floats in scan order; a
// parallel sum cannot reproduce that rounding, so partial sums are kept in double and combined in a fixed order.
__global__ void __launch_bounds__(256) k_luma_partial(const float4 *px, uint32_t n, double *partial) {
    __shared__ double sh[8];
    double s = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float l = color_luma(px[i]);
        if (l > 0.0f) s += (double)logf(0.01f + l);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += sh[w]; partial[blockIdx.x] = t; }
}
__global__ void k_luma_final(const double *partial, uint32_t nb, uint32_t n_pixels, float *scene_luma) {
    if (blockIdx.x || threadIdx.x) return;
    double t = 0.0;
    for (uint32_t i = 0; i < nb; ++i) t += partial[i];
    float lavg = (float)t;
    *scene_luma = expf(lavg / (float)n_pixels);                                                  // main.cpp:98
}
__global__ void k_tonemap_pack(const float4 *px, uint32_t n, const float *scene_luma_ptr, uchar4 *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 c = px[i];
    float scene_luma = *scene_luma_ptr;
    float key_alpha = 0.18f;                                                                     // main.cpp:115-123
    float pixel_luma = color_luma(c);
    float l_xy = key_alpha * pixel_luma / scene_luma;
    float l_d = l_xy / (1.0f + l_xy);
    float scale = l_d / pixel_luma;
    c.x *= scale; c.y *= scale; c.z *= scale;
    uchar4 o;                                                                                    // Color_Pack, color.h:105-111
    o.x = (unsigned char)(clampf(c.x, 0.0f, 1.0f) * 255.0f);
    o.y = (unsigned char)(clampf(c.y, 0.0f, 1.0f) * 255.0f);
    o.z = (unsigned char)(clampf(c.z, 0.0f, 1.0f) * 255.0f);
    o.w = (unsigned char)(clampf(c.w, 0.0f, 1.0f) * 255.0f);
    out[i] = o;
}
