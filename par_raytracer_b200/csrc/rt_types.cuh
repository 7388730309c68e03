// rt_types.cuh -- plain structs shared by the kernels (rt_trace.cuh, rt_shade.cuh) and the host translation units
// (rt_internal.h): ray / hit / shadow queues, the per-path pool, the wave descriptor. No kernels here, so every TU may include it.
#pragma once
#include "rt_common.cuh"
#include "rt_raygen.cuh"

struct RayQueue {              // SoA ray stream: 32 B per ray
    float4 *o;                 // origin.xyz (unbiased, as handed to TraceRay), w: kernel specific
    float4 *d;                 // direction.xyz, w: slot / flags bits
};

struct HitRec {                // 16 B per ray
    float t;
    float v, w;                // bw.y, bw.z numerators already divided (raytracer.cpp:118-119)
    int32_t tri;               // cluster-order triangle index, -1 = miss
};

struct TraceCounters { unsigned long long sphere_checks, cluster_checks; };

struct PathPool {
    uint4 *rng_cx;               // (cur.lo, cur.hi, x.lo, x.hi) of the 28-byte generator state (rt_rng.cuh)
    uint64_t *rng_seed;          // only read on the rare > 15-draws replay path
    float4 *acc;                 // xyz: radiance gathered by the sample so far
    float4 *node_T;              // xyz: throughput of the in-flight node, w: iters | frames << 8 | draws << 16
    float4 *frames;              // [(level * RT_FRAME_F4 + k) * capacity + slot]
    uint32_t *ray_cnt;           // NULL, or per path: TraceRay calls of this sample so far (adaptive sampling keeps the counts of discarded samples out of ray_count)
    uint32_t capacity;
};

struct ShadowQueue { float4 *o; float4 *rad; uint32_t *count; uint32_t capacity; };   // o: origin.xyz + slot; rad: radiance.xyz + light_dist_sq (< 0: directional)

struct WaveQueues {
    RayQueue closest;              // d.w = path slot
    const uint32_t *n_closest;     // device-side count (NULL: closest_max rays)
    uint32_t closest_max;
    HitRec *hits;
    const float4 *shadow_o;        // light l owns [l * shadow_stride, ...): origin.xyz, w = path slot
    const float4 *shadow_dir;      // NULL: direction = f(light, origin) as GetShadowRayForLight (raytracer.cpp:234-250); else explicit
    const float4 *rad;             // radiance to add when the light is visible; w = light_dist_sq (point light) or < 0
    const uint32_t *n_shadow;      // [n_lights] device-side counts
    uint32_t shadow_stride, n_lights;
    float4 *acc;                   // light 0 adds into the path accumulator ...
    float4 *acc_extra;             // ... light l >= 1 into acc_extra[(l - 1) * shadow_stride + slot] (single writer each: no atomics)
    uint32_t *next;                // work counter, zero before launch
    uint32_t fetch_min;            // refill the warp once this many lanes are idle (32 = only when all are): bounce rays
    uint32_t fetch_min_primary;    // same while the work counter is still inside the primary rays of wave 0
    uint32_t fetch_min_shadow;     // same inside the shadow-ray region
    uint32_t leaf_wait;            // leave the node loop once this many live lanes wait at a cluster / have finished (32: only when all do)
};
