// rt_raygen.cuh -- K1: primary ray generation = RenderPixel's per-sample prologue (main.cpp:237-241 / 246-250) and
// MakeCameraRay (main.cpp:164-177), bit-exact. The wave-0 trace and shading kernels call primary_ray() directly for
// slot s instead of streaming a ray queue and an initial path state through HBM (~200 B per primary sample).
#pragma once
#include "rt_common.cuh"
#include "rt_rng.cuh"

struct DevCamera { float tan_a2, aspect, inv_width, inv_height; float pos[3], fwd[3], right[3], up[3]; };   // == Camera, main.cpp:133-143

RT_DEVICE void camera_ray(const DevCamera &cam, float ox, float oy, f3 &org, f3 &dir) {        // main.cpp:164-177
    float nx = 2.0f * (ox + 0.5f) * cam.inv_width - 1.0f;
    float ny = 1.0f - 2.0f * (oy + 0.5f) * cam.inv_height;
    f3 fwd = mk3(cam.fwd[0], cam.fwd[1], cam.fwd[2]);
    f3 right = mk3(cam.right[0], cam.right[1], cam.right[2]);
    f3 up = mk3(cam.up[0], cam.up[1], cam.up[2]);
    f3 a = ((right * cam.tan_a2) * cam.aspect) * nx;
    f3 b = (up * cam.tan_a2) * ny;
    dir = normalize3((fwd + a) + b);
    org = mk3(cam.pos[0], cam.pos[1], cam.pos[2]);
}

// slot s of a batch <-> (pixel_local = pixel_local0 + s / spp, sample = sample_begin + s % spp)
struct PrimaryGen {
    DevCamera cam;
    uint64_t base_seed;
    const uint32_t *pixel_ids;      // device; NULL: linear range starting at pixel_begin
    uint32_t n_slots, spp, width, pixel_begin, pixel_local0, sample_begin;
    float jitter_scale;             // 0.5 in RenderPixel's first loop, 1.0 in its adaptive loop (main.cpp:240 vs 249)
    uint32_t enabled;               // 0: rays and path state come from the queues
};

// generator state of slot s after its two jitter draws (what primary_ray leaves in r)
RT_DEVICE void primary_rng(const PrimaryGen &g, uint32_t s, PathRng &r) {
    uint32_t pl = g.pixel_local0 + s / g.spp;
    uint32_t samp = g.sample_begin + s % g.spp;
    uint32_t pixel = g.pixel_ids ? g.pixel_ids[pl] : g.pixel_begin + pl;
    rng_seed(r, sample_seed(g.base_seed, pixel, samp));
    (void)rng_next(r);
    (void)rng_next(r);
}

RT_DEVICE void primary_ray(const PrimaryGen &g, uint32_t s, PathRng &r, f3 &org, f3 &dir) {
    uint32_t pl = g.pixel_local0 + s / g.spp;
    uint32_t samp = g.sample_begin + s % g.spp;
    uint32_t pixel = g.pixel_ids ? g.pixel_ids[pl] : g.pixel_begin + pl;
    uint32_t x = pixel % g.width, y = pixel / g.width;       // main.cpp:274-275
    rng_seed(r, sample_seed(g.base_seed, pixel, samp));
    float jy = rng_float11(r);                               // y takes the first draw
    float jx = rng_float11(r);
    camera_ray(g.cam, (float)x + jx * g.jitter_scale, (float)y + jy * g.jitter_scale, org, dir);
}
