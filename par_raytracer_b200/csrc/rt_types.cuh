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

// Layout of the per-path records (RT_POOL_AOS, default 1). Bounce waves touch slots in no particular order, so what a thread moves in one go should
// sit in whole 32-byte sectors of its own: the in-flight node's throughput and the generator state are ONE 32-byte record per slot (two scattered
// half-used sectors otherwise), and the words of a recursion frame are contiguous per (slot, level), the four words the reference's default 1 + 1
// samples use first (64 bytes = two full sectors instead of four half-used ones). 0 = the round-1 structure-of-arrays form (one array per word).
#ifndef RT_POOL_AOS
#define RT_POOL_AOS 1
#endif
#define RT_FRAME_F4 6            // float4 words per recursion frame

struct PathPool {
    uint4 *rng_cx;               // (cur.lo, cur.hi, x.lo, x.hi) of the 28-byte generator state (rt_rng.cuh); RT_POOL_AOS: unused, the state lives in node_T[2 * slot + 1]
    uint64_t *rng_seed;          // only read on the rare > 15-draws replay path
    float4 *acc;                 // xyz: radiance gathered by the sample so far
    float4 *node_T;              // xyz: throughput of the in-flight node, w: iters | frames << 8 | draws << 16; RT_POOL_AOS: node_T[2 * slot]
    float4 *frames;              // RT_POOL_AOS: [(slot * depth + level) * RT_FRAME_F4 + word]; else [(level * RT_FRAME_F4 + word) * capacity + slot]
    uint32_t *ray_cnt;           // NULL, or per path: TraceRay calls of this sample so far (adaptive sampling keeps the counts of discarded samples out of ray_count)
    uint32_t capacity;
    uint32_t depth;              // recursion frames per slot
};
#if RT_POOL_AOS
RT_DEVICE float4 *path_T(const PathPool &P, uint32_t slot) { return P.node_T + 2 * (size_t)slot; }
RT_DEVICE uint4 *path_rng(const PathPool &P, uint32_t slot) { return reinterpret_cast<uint4 *>(P.node_T) + 2 * (size_t)slot + 1; }
// logical frame words: 0 position + meta, 1 normal + material, 2 incoming direction, 3 diffuse weight, 4 specular weight, 5 continuation origin;
// stored in the order 0 1 2 4 3 5
RT_DEVICE float4 *frame_word(const PathPool &P, uint32_t slot, uint32_t level, uint32_t k) {
    const uint32_t phys = k == 3u ? 4u : (k == 4u ? 3u : k);
    return P.frames + ((size_t)slot * P.depth + level) * RT_FRAME_F4 + phys;
}
#else
RT_DEVICE float4 *path_T(const PathPool &P, uint32_t slot) { return P.node_T + slot; }
RT_DEVICE uint4 *path_rng(const PathPool &P, uint32_t slot) { return P.rng_cx + slot; }
RT_DEVICE float4 *frame_word(const PathPool &P, uint32_t slot, uint32_t level, uint32_t k) {
    return P.frames + (size_t)(level * RT_FRAME_F4 + k) * P.capacity + slot;
}
#endif

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
