// rt_shade.cuh -- kernel groups K1 (ray generation), K4/K5 (shading, texture sampling, bounce generation),
// K6 (stream compaction between stages) and K7 (accumulate / resolve).
//
// TraceRayColor (raytracer.cpp:413-577) is a recursion TREE whose nodes draw from ONE sequential random
// stream in depth-first order, and how many draws a subtree takes depends on what its rays hit. To keep
// every draw of every (pixel, sample) identical to the reference's, each sample is run as a coroutine
// that executes the reference's recursion in the reference's order and yields at every TraceRay call:
//
//   wave w:   k_trace_closest  : all pending nodes' rays (one per live sample)        -> HitRec stream
//             k_logic          : per live sample: consume the hit, shade, queue the shadow rays of the
//                                node (they take no random numbers, so they ride in the same wave), then
//                                walk the recursion forward -- draw, Russian-roulette, pick the next child
//                                in DFS order -- until the next ray that really has to be traced; that ray
//                                is appended (warp-aggregated atomics) to the next wave's compacted queue
//             k_trace_shadow   : occlusion of the queued shadow rays, adds the pre-weighted radiance
//
// Radiance flows forward as a throughput T (component-wise products of the reference's own factors:
// alpha, diffuse_color * w_diffuse * max(0, N.dir), specular_color * max(0, dir.-V), 1 - alpha), which is the
// reference's nested sum re-associated; everything that steers control flow is bit-identical.
#pragma once
#include "rt_common.cuh"
#include "rt_rng.cuh"
#include "rt_raygen.cuh"
#include "rt_trace.cuh"
#include "rt_types.cuh"



RT_DEVICE uint32_t pack_state(int iters, uint32_t sp, uint32_t draws) { return (uint32_t)iters | (sp << 8) | (draws << 16); }

RT_DEVICE void rng_unpack(uint4 cx, uint32_t draws, PathRng &r) {
    r.cur = (uint64_t)cx.x | ((uint64_t)cx.y << 32); r.x = (uint64_t)cx.z | ((uint64_t)cx.w << 32); r.n = draws; r.seed = 0;
}
RT_DEVICE uint4 rng_pack(const PathRng &r) {
    return make_uint4((uint32_t)r.cur, (uint32_t)(r.cur >> 32), (uint32_t)r.x, (uint32_t)(r.x >> 32));
}

// One atomicAdd per warp; every lane of the warp must call it.
RT_DEVICE uint32_t warp_push(uint32_t *counter, bool pred) {
    uint32_t mask = __ballot_sync(0xffffffffu, pred);
    uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == 0 && mask) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// ---- K1: ray generation = RenderPixel's per-sample prologue (main.cpp:237-241 / 246-250) ---------------
// slot s of the batch <-> (pixel_local = s / spp, sample = sample_begin + s % spp)
__global__ void k_raygen(PrimaryGen g, DevParams prm, PathPool P, RayQueue q, uint32_t *n_rays_out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n_slots) return;
    PathRng r;
    f3 org, dir;
    primary_ray(g, s, r, org, dir);
    q.o[s] = mk4(org, 0.0f);
    q.d[s] = mk4u(dir, s);
    P.rng_seed[s] = r.seed;
    *path_rng(P, s) = rng_pack(r);
    P.acc[s] = make_float4(0, 0, 0, 0);
    *path_T(P, s) = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pack_state((int)prm.bounce_depth, 0, r.n)));
    if (s == 0) *n_rays_out = g.n_slots;
}

// paths started from caller-supplied rays and seeds (rt_trace_color <-> TraceRayColor)
__global__ void k_paths_from_rays(DevParams prm, PathPool P, RayQueue q, uint32_t n, const float *rays, const uint64_t *seeds,
                                  uint32_t *n_rays_out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float *rp = rays + 6 * (size_t)s;
    q.o[s] = make_float4(rp[0], rp[1], rp[2], 0.0f);
    q.d[s] = make_float4(rp[3], rp[4], rp[5], __uint_as_float(s));
    PathRng r;
    rng_seed(r, seeds[s]);
    P.rng_seed[s] = r.seed;
    *path_rng(P, s) = rng_pack(r);
    P.acc[s] = make_float4(0, 0, 0, 0);
    *path_T(P, s) = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pack_state((int)prm.bounce_depth, 0, r.n)));
    if (s == 0) *n_rays_out = n;
}

// ---- K5: texture.cpp:5-83 --------------------------------------------------------------------------
RT_DEVICE float wrap_uv(float uv) { return uv >= 0.0f ? fmodf(uv, 1.0f) : 1.0f + fmodf(uv, 1.0f); }

RT_DEVICE f3 get_texel(const DevScene &S, const DevTexture &t, uint32_t x, uint32_t y) {      // texture.cpp:17-51 (xyz only)
    const uint8_t *p = S.texels + t.offset + (size_t)(y * t.size_x + x) * t.channels;
    uint32_t r = p[0], g = 0, b = 0;
    if (t.channels >= 2) g = p[1];
    if (t.channels >= 3) b = p[2];
    if (t.channels == 1) { g = r; b = r; }
    return mk3(__ldg(S.srgb_lut + r), __ldg(S.srgb_lut + g), __ldg(S.srgb_lut + b));
}
RT_DEVICE f3 lerp3(f3 a, f3 b, float t) { return a + (b - a) * t; }                           // mathlib.h:10

RT_DEVICE f3 tex_sample(const DevScene &S, int tex, float u, float v) {                        // texture.cpp:53-83
    DevTexture t = S.textures[tex];
    u = wrap_uv(u);
    v = 1.0f - wrap_uv(v);
    float sx = (float)(t.size_x - 2u);
    float sy = (float)(t.size_y - 2u);
    float tx = clampf(u * sx, 0.0f, sx);
    float ty = clampf(v * sy, 0.0f, sy);
    uint32_t tx0 = (uint32_t)floorf(tx), ty0 = (uint32_t)floorf(ty);
    uint32_t tx1 = tx0 + 1u, ty1 = ty0 + 1u;
    float fx = tx - (float)tx0;
    float fy = ty - (float)ty0;
    f3 s00 = get_texel(S, t, tx0, ty0);
    f3 s01 = get_texel(S, t, tx0, ty1);
    f3 s10 = get_texel(S, t, tx1, ty0);
    f3 s11 = get_texel(S, t, tx1, ty1);
    return lerp3(lerp3(s00, s01, fy), lerp3(s10, s11, fy), fx);
}

// ---- raytracer.cpp:302-371 ---------------------------------------------------------------------------
RT_DEVICE f3 to_world(f3 normal, f3 local) {                                                   // raytracer.cpp:306-312
    f3 up = fabsf(normal.z) < 0.9999f ? mk3(0.0f, 0.0f, 1.0f) : mk3(1.0f, 0.0f, 0.0f);
    f3 tangent = normalize3(cross3(up, normal));
    f3 bitangent = normalize3(cross3(normal, tangent));
    return normalize3((tangent * local.x + bitangent * local.y) + normal * local.z);
}
RT_DEVICE f3 reflect3(f3 v, f3 n) { return ((n * 2.0f) * dot3(v, n)) - v; }                    // raytracer.cpp:343-346
RT_DEVICE float fresnel_amount(float ior_exit, float ior_enter, f3 normal, f3 incident) {      // raytracer.cpp:348-371
    float r0 = (ior_exit - ior_enter) / (ior_exit + ior_enter);
    r0 *= r0;
    float ct = max0(-dot3(normal, incident));
    if (ior_exit > ior_enter) {
        float n = ior_exit / ior_enter;
        float st_sq = n * n * (1.0f - ct * ct);
        if (st_sq > 1.0f) return 1.0f;
        ct = sqrtf(1.0f - st_sq);
    }
    float x = 1.0f - ct;
    float x2 = x * x;
    float x3 = x * x2;
    return r0 + (1.0f - r0) * x2 * x3;
}
// powf(Max(0, spec_cos), Ns) (raytracer.cpp:388, 403). glibc's powf cannot be reproduced bit-for-bit on the GPU; CUDA's
// powf is within a few ulp of it and the value only scales a colour (never control flow). -DRT_PHONG_POW_F64 switches to
// the double-precision pow rounded to float (<= 1 ulp from glibc, ~4x the instructions).
#ifdef RT_PHONG_POW_F64
RT_DEVICE float phong_pow(float x, float e) { return (float)pow((double)x, (double)e); }
#else
RT_DEVICE float phong_pow(float x, float e) { return (x == 0.0f && e > 0.0f) ? 0.0f : powf(x, e); }   // pow(+0, e > 0) = +0: skip the call
#endif

// One sub-queue per light (light l owns entries [l * capacity, (l + 1) * capacity) and count[l]): a path has at most one
// ray in each, so the occlusion kernel adds to the path's accumulator without atomics and in light order.

// ---- K4 + K6: the coroutine step --------------------------------------------------------------------
#ifndef RT_LOGIC_MIN_BLOCKS
#define RT_LOGIC_MIN_BLOCKS 6
#endif
#ifndef RT_FRAME_LAST_POP
#define RT_FRAME_LAST_POP 1      // 1: a recursion frame is popped when its last child is taken and holds only the words later children read (0: every frame is pushed whole and re-read once exhausted)
#endif
#ifndef RT_LOGIC_CHUNK
#define RT_LOGIC_CHUNK 256       // queue entries a block sorts by material and deals to its 128 threads at a time (multiple of 128, <= 65536; measured 256 / 512 / 1024 / 2048: profiles/README.md)
#endif
__global__ void __launch_bounds__(128, RT_LOGIC_MIN_BLOCKS) k_logic(DevScene S, DevParams prm, PathPool P, RayQueue qin, const HitRec *hits, const uint32_t *n_in_ptr,
                                              uint32_t n_in_max, RayQueue qout, uint32_t *n_out, ShadowQueue sh, PrimaryGen G) {
    const uint32_t n_in = G.enabled ? G.n_slots : min(*n_in_ptr, n_in_max);
    const int bd = (int)prm.bounce_depth;
    // Persistent blocks, block-uniform trip counts (every lane of a warp reaches the warp-aggregated pushes together). A block takes
    // RT_LOGIC_CHUNK consecutive queue entries at a time and deals them to its threads SORTED by (miss | material of the hit triangle):
    // a counting sort of the chunk's hit records in shared memory. Bounce rays of neighbouring queue entries hit unrelated materials or
    // the sky, and the shading code of a textured / bump-mapped / masked material differs from the plain one's: unsorted, ncu showed 10 of
    // 32 lanes per instruction (4 in the texture code) on the waves after the first; after the deal a warp mostly holds one kind of work.
    // The order inside a key is arrival order -- irrelevant: paths are independent, and every path owns its slot.
    __shared__ HitRec s_hit[RT_LOGIC_CHUNK];
    __shared__ uint16_t s_perm[RT_LOGIC_CHUNK];
    __shared__ uint32_t s_bin[64];
    for (uint32_t chunk = blockIdx.x * RT_LOGIC_CHUNK; chunk < n_in; chunk += gridDim.x * RT_LOGIC_CHUNK) {
    const uint32_t n_chunk = min((uint32_t)RT_LOGIC_CHUNK, n_in - chunk);
    {
        constexpr int PER = RT_LOGIC_CHUNK / 128;
        uint32_t my_key[PER], my_rank[PER];
        if (threadIdx.x < 64) s_bin[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const uint32_t e = threadIdx.x + 128u * j;
            my_key[j] = 0; my_rank[j] = 0;
            if (e < n_chunk) {
                const HitRec h = hits[chunk + e];
                s_hit[e] = h;
                my_key[j] = h.tri < 0 ? 0u : 1u + (uint32_t)(__ldg(S.tri_mat + h.tri) % 63u);
                my_rank[j] = atomicAdd(&s_bin[my_key[j]], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {              // exclusive prefix over the 64 bins
            const uint32_t a = s_bin[2 * threadIdx.x], b = s_bin[2 * threadIdx.x + 1];
            uint32_t incl = a + b;
            for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += v; }
            s_bin[2 * threadIdx.x] = incl - a - b; s_bin[2 * threadIdx.x + 1] = incl - b;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const uint32_t e = threadIdx.x + 128u * j;
            if (e < n_chunk) s_perm[s_bin[my_key[j]] + my_rank[j]] = (uint16_t)e;
        }
        __syncthreads();
    }
    for (uint32_t trip = 0; trip < RT_LOGIC_CHUNK; trip += 128) {
    if (trip >= n_chunk) break;                     // block-uniform
    const uint32_t pidx = trip + threadIdx.x;
    const bool active = pidx < n_chunk;
    const uint32_t e_local = active ? (uint32_t)s_perm[pidx] : 0u;
    const uint32_t i = chunk + e_local;
    HitRec h_cur; h_cur.t = 0.0f; h_cur.v = 0.0f; h_cur.w = 0.0f; h_cur.tri = -1;
    if (active) h_cur = s_hit[e_local];

    uint32_t slot = 0, sp = 0;
    int iters = 0;
    f3 T = mk3(0, 0, 0), acc = mk3(0, 0, 0);
    PathRng rng; rng.cur = 0; rng.x = 0; rng.seed = 0; rng.n = 0;
    bool emit = false;
    f3 e_org = mk3(0, 0, 0), e_dir = mk3(0, 0, 0), e_T = mk3(0, 0, 0);
    int e_iters = 0;
    bool shade_hit = false, fresh_frame = false;      // fresh_frame: the top frame was pushed by this thread, its fields are still in registers
    f3 f_Td = mk3(0, 0, 0);
    f3 hit_p = mk3(0, 0, 0), N = mk3(0, 0, 0), V = mk3(0, 0, 0), Ta = mk3(0, 0, 0), kd = mk3(0, 0, 0), ks = mk3(0, 0, 0);
    float spec_int = 0.0f, w_diffuse = 0.0f;

    if (active) {
        const HitRec h = h_cur;
        f3 org;
        if (G.enabled) {
            // wave 0: the primary ray and the fresh path state are recomputed, not streamed (rt_raygen.cuh); a primary
            // miss needs neither (the path ends with the background colour)
            slot = i;
            if (h.tri >= 0) {
                // the trace kernel left the direction it generated in the queue (hits only); the origin is the camera's;
                // the generator state after the two jitter draws is re-derived (cheaper than streaming it)
                V = mk3(qin.d[i]);
                org = mk3(G.cam.pos[0], G.cam.pos[1], G.cam.pos[2]);
                primary_rng(G, i, rng);
            } else org = mk3(0, 0, 0);
            T = mk3(1.0f, 1.0f, 1.0f); iters = bd; sp = 0;
        } else {
            // every independent load of the path's state is issued before the first use
            float4 o4 = qin.o[i], d4 = qin.d[i];
            slot = __float_as_uint(d4.w);
            float4 t4 = *path_T(P, slot);
            uint4 cx = *path_rng(P, slot);
            float4 a4 = P.acc[slot];
            uint32_t st = __float_as_uint(t4.w);
            T = mk3(t4); iters = (int)(st & 0xFFu); sp = (st >> 8) & 0xFFu;
            rng_unpack(cx, st >> 16, rng);
            if (rng.n >= 15u) rng.seed = P.rng_seed[slot];
            acc = mk3(a4);
            org = mk3(o4); V = mk3(d4);
        }

        if (h.tri < 0) {
            acc = acc + T * mk3(prm.bg[0], prm.bg[1], prm.bg[2]);          // raytracer.cpp:573-575
        } else {
            float4 ua = __ldg(S.tri_uv + 2 * (size_t)h.tri), ub = __ldg(S.tri_uv + 2 * (size_t)h.tri + 1);
            const float4 *np = S.tri_nrm + 3 * (size_t)h.tri;
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            f3 ob = org + V * prm.ray_bias;                                  // raytracer.cpp:163
            f3 position = ob + V * h.t;                                      // raytracer.cpp:121
            f3 gn = mk3(n0.w, n1.w, n2.w);                                   // raytracer.cpp:122, normalised once at build time
            hit_p = position + gn * prm.ray_bias;                            // raytracer.cpp:425
            int mat_id = __float_as_int(ub.z);
            DevMaterial M = S.materials[mat_id];
            float bwy = h.v, bwz = h.w;
            float bwx = 1.0f - bwy - bwz;                                    // raytracer.cpp:120
            float u = 0.0f + ua.x * bwx; u = u + ua.z * bwy; u = u + ub.x * bwz;   // raytracer.cpp:439-442
            float v = 0.0f + ua.y * bwx; v = v + ua.w * bwy; v = v + ub.y * bwz;
            float alpha = M.alpha;
            bool restart = false;
            if (M.alpha <= 1.0f || M.tex_alpha >= 0) {                       // raytracer.cpp:443-453
                if (M.tex_alpha >= 0) alpha *= tex_sample(S, M.tex_alpha, u, v).x;
                if (alpha <= 0.05f) restart = true;
            }
            if (restart) {
                // tail call TraceRayColor(ray, iters): its Russian roulette runs again unless this is the root level
                if (iters != bd && rng.n >= 15u && rng.seed == 0) rng.seed = P.rng_seed[slot];
                if (!(iters != bd && rng_float01(rng) < 0.5f)) {
                    emit = true;
                    e_org = position + (V * prm.ray_bias) * 2.0f;            // raytracer.cpp:450
                    e_dir = V; e_T = T; e_iters = iters;
                }
            } else {
                f3 ambient = mk3(M.ambient[0], M.ambient[1], M.ambient[2]);
                kd = mk3(M.diffuse[0], M.diffuse[1], M.diffuse[2]);
                ks = mk3(M.specular[0], M.specular[1], M.specular[2]);
                if (M.tex_ambient >= 0) ambient = ambient * tex_sample(S, M.tex_ambient, u, v);      // raytracer.cpp:454-462
                if (M.tex_diffuse >= 0) kd = kd * tex_sample(S, M.tex_diffuse, u, v);
                if (M.tex_specular >= 0) ks = tex_sample(S, M.tex_specular, u, v);
                f3 n = mk3(0, 0, 0);                                          // raytracer.cpp:464-467
                n = n + mk3(n0) * bwx; n = n + mk3(n1) * bwy; n = n + mk3(n2) * bwz;
                N = normalize3(n);
                if (M.tex_bump >= 0) {                                       // raytracer.cpp:468-502
                    const float4 *gp = S.tri_tan + 3 * (size_t)h.tri;
                    f3 tg = mk3(0, 0, 0);
                    tg = tg + mk3(__ldg(gp)) * bwx; tg = tg + mk3(__ldg(gp + 1)) * bwy; tg = tg + mk3(__ldg(gp + 2)) * bwz;
                    tg = normalize3(tg);
                    f3 bt = normalize3(cross3(N, tg));
                    f3 sn = tex_sample(S, M.tex_bump, u, v) * 2.0f - mk3(1.0f, 1.0f, 1.0f);
                    f3 wn;
                    wn.x = tg.x * sn.x + bt.x * sn.y + N.x * sn.z;
                    wn.y = tg.y * sn.x + bt.y * sn.y + N.y * sn.z;
                    wn.z = tg.z * sn.x + bt.z * sn.y + N.z * sn.z;
                    N = wn;                                                  // not re-normalised (raytracer.cpp:494-495)
                }
                spec_int = M.specular_intensity;
                float object_reflectivity = 0.04f;                           // raytracer.cpp:538-541
                float fres = fresnel_amount(1.0f, M.index_of_refraction, N, V);
                float w_reflect = (object_reflectivity + (1.0f - object_reflectivity) * fres);
                w_diffuse = 1.0f - w_reflect;
                bool translucent = alpha < 1.0f;                             // raytracer.cpp:547-552
                Ta = translucent ? T * alpha : T;
                acc = acc + Ta * (ambient * 0.1f);                            // raytracer.cpp:543
                shade_hit = true;
                if (iters > 0) {
                    // Children exist. The first one (a diffuse sample, if there are any) is generated from registers further down; the node's
                    // recursion frame is pushed only if more children follow, and only with the words those children read: F[2] / F[4] (V,
                    // specular weight) for specular children and the translucent continuation, F[3] (diffuse weight) for diffuse children
                    // after the first, F[5] for the continuation. Reference defaults (1 + 1 samples, opaque): 64 bytes, read once.
                    const uint32_t rs = prm.reflection_samples, ss = prm.spec_samples;
                    f3 Td = (Ta * kd) * w_diffuse;                           // raytracer.cpp:544
                    f3 Ts = Ta * ks;                                         // raytracer.cpp:545
                    f3 Tc = T * (1.0f - alpha);
                    fresh_frame = rs > 0;
                    f_Td = Td;
                    const bool lean = RT_FRAME_LAST_POP != 0;
                    if (!lean || rs + ss + (translucent ? 1u : 0u) > (fresh_frame ? 1u : 0u)) {
                        uint32_t meta = (uint32_t)iters | (translucent ? 0x80000000u : 0u) | (fresh_frame ? (1u << 8) : 0u);   // next child index in bits 8..30
                        *frame_word(P, slot, sp, 0) = mk4u(hit_p, meta);
                        *frame_word(P, slot, sp, 1) = mk4u(N, (uint32_t)mat_id);
                        if (!lean || ss > 0 || translucent) {
                            *frame_word(P, slot, sp, 2) = mk4(V, spec_int);
                            *frame_word(P, slot, sp, 4) = mk4(Ts, Tc.y);
                        }
                        if (!lean || rs > 1 || translucent) *frame_word(P, slot, sp, 3) = mk4(Td, Tc.x);
                        if (translucent) *frame_word(P, slot, sp, 5) = mk4(position + (V * prm.ray_bias) * 2.0f, Tc.z);   // raytracer.cpp:549
                        sp++;
                    }
                }
                // iters == 0: the translucent continuation has iters - 1 < 0 and returns black (raytracer.cpp:416)
            }
        }
    }

    // ShadeLight (raytracer.cpp:378-411): everything but the visibility test is evaluated here; the shadow ray carries the
    // pre-weighted radiance. Its direction is a function of (light, origin) and is re-derived by the trace kernel.
    for (uint32_t l = 0; l < S.n_lights; ++l) {
        f3 rad = mk3(0, 0, 0);
        float dist_sq = -1.0f;
        const uint32_t spos = warp_push(sh.count + l, shade_hit) + l * sh.capacity;     // issued first: the atomic's latency hides under the shading math
        if (shade_hit) {
            DevLight L = S.lights[l];
            f3 lc = mk3(L.color[0], L.color[1], L.color[2]);
            f3 lv;
            if (L.type == 0) {
                lv = mk3(L.facing[0], L.facing[1], L.facing[2]) * -1.0f;     // raytracer.cpp:240
            } else {
                f3 lp = mk3(L.position[0], L.position[1], L.position[2]);
                lv = normalize3(lp - hit_p);                                 // raytracer.cpp:243
                f3 dv = lp - hit_p;
                dist_sq = dot3(dv, dv);
                float fd = (sqrtf(dist_sq) / L.falloff) + 1.0f;              // raytracer.cpp:398-399
                lc = lc * (1.0f / (fd * fd));
            }
            float spec_cos = dot3(V * -1.0f, reflect3(lv, N));               // raytracer.cpp:386-388
            f3 dd = (lc * 2.0f) * max0(dot3(N, lv));
            f3 ds = lc * phong_pow(max0(spec_cos), spec_int);
            rad = Ta * (((dd * kd) * w_diffuse) + ds * ks);                  // raytracer.cpp:544-545
        }
        if (shade_hit) {
            sh.o[spos] = mk4u(hit_p, slot);
            sh.rad[spos] = mk4(rad, dist_sq);
        }
    }

    if (active) {
        // walk the recursion forward to the next ray that has to be traced
        if (fresh_frame) {
            // first child of the node shaded above (a diffuse sample, raytracer.cpp:516-526): same arithmetic as the generic
            // walk below, operands taken from registers instead of re-reading the frame just written (its child index
            // was stored as 1 already)
            if (rng.n >= 14u && rng.seed == 0) rng.seed = P.rng_seed[slot];
            uint32_t series_i = (uint32_t)(rng_next(rng) % 1024ull);
            f3 c_dir = to_world(N, mk3(__ldg(S.hamm_dir + series_i)));
            f3 c_T = f_Td * max0(dot3(N, c_dir));
            if (!(rng_float01(rng) < 0.5f)) { emit = true; e_org = hit_p; e_dir = c_dir; e_T = c_T; e_iters = iters - 1; }
        }
        while (!emit && sp > 0) {
            const uint32_t lvl = sp - 1;
            float4 *F = frame_word(P, slot, lvl, 0);
            float4 f0 = F[0];
            uint32_t meta = __float_as_uint(f0.w);
            uint32_t child = (meta >> 8) & 0x7FFFFFu;
            int f_iters = (int)(meta & 0xFFu);
            bool has_cont = (meta & 0x80000000u) != 0;
            uint32_t rs = prm.reflection_samples, ss = prm.spec_samples;
            const uint32_t n_children = rs + ss + (has_cont ? 1u : 0u);
            if (child >= n_children) { sp--; continue; }                      // RT_FRAME_LAST_POP: guard only, a frame leaves the stack with its last child
            // The LAST child pops the frame as it is taken (its words are read below, nothing overwrites them before): a path that comes
            // back from that child finds the grandparent on top and never re-reads an exhausted frame -- one dependent load less per
            // return, no child-index store for the frames that have one child left (all of them with the reference's 1 + 1 samples).
            if (RT_FRAME_LAST_POP && child + 1u >= n_children) sp--;
            else F[0].w = __uint_as_float((meta & 0x800000FFu) | ((child + 1u) << 8));
            if (rng.n >= 14u && rng.seed == 0) rng.seed = P.rng_seed[slot];
            f3 fp = mk3(f0);
            float4 f1 = *frame_word(P, slot, lvl, 1);
            f3 fn = mk3(f1);
            f3 c_dir, c_T, c_org = fp;
            if (child < rs) {                                                // raytracer.cpp:516-526
                uint32_t series_i = (uint32_t)(rng_next(rng) % 1024ull);
                c_dir = to_world(fn, mk3(__ldg(S.hamm_dir + series_i)));
                float4 f3v = *frame_word(P, slot, lvl, 3);
                c_T = mk3(f3v) * max0(dot3(fn, c_dir));
            } else if (child < rs + ss) {                                    // raytracer.cpp:528-535
                uint32_t mat_id = __float_as_uint(f1.w);
                c_dir = to_world(fn, mk3(__ldg(S.spec_dir + (size_t)mat_id * ss + (child - rs))));
                float4 f2 = *frame_word(P, slot, lvl, 2), f4 = *frame_word(P, slot, lvl, 4);
                c_T = mk3(f4) * max0(dot3(c_dir, neg3(mk3(f2))));
            } else {                                                         // raytracer.cpp:547-551
                float4 f2 = *frame_word(P, slot, lvl, 2), f3v = *frame_word(P, slot, lvl, 3), f4 = *frame_word(P, slot, lvl, 4), f5 = *frame_word(P, slot, lvl, 5);
                c_dir = mk3(f2); c_org = mk3(f5); c_T = mk3(f3v.w, f4.w, f5.w);
            }
            // child node entry (raytracer.cpp:416-420): iters - 1 >= 0 here, and never the root level
            if (rng_float01(rng) < 0.5f) continue;
            emit = true; e_org = c_org; e_dir = c_dir; e_T = c_T; e_iters = f_iters - 1;
        }
    }
    // K6: compaction -- the next wave's queue holds only rays that exist (the atomic is issued before the state stores)
    uint32_t pos = warp_push(n_out, emit);
    if (active) {
        P.acc[slot] = mk4(acc, 0.0f);
        if (P.ray_cnt) {                               // rays this sample has asked for: its primary ray, every emitted ray, one shadow ray per light and shade
            const uint32_t add = (emit ? 1u : 0u) + (shade_hit ? S.n_lights : 0u);
            if (G.enabled) P.ray_cnt[slot] = 1u + add;
            else if (add) P.ray_cnt[slot] += add;
        }
        if (emit) {                                    // a finished path only leaves its radiance behind
            *path_rng(P, slot) = rng_pack(rng);
            *path_T(P, slot) = mk4u(e_T, pack_state(e_iters, sp, rng.n));
            if (G.enabled) P.rng_seed[slot] = rng.seed;
        }
    }
    if (emit) { qout.o[pos] = mk4(e_org, 0.0f); qout.d[pos] = mk4u(e_dir, slot); }
    }
    __syncthreads();                                // the next chunk reuses the shared arrays
    }
}

// ---- K7: accumulate / resolve (main.cpp:242, 262-263) -----------------------------------------------
// accum[p] continues RenderPixel's `color += scratch[samp]` in sample order across sample sub-batches.
__global__ void k_resolve(const float4 *acc, uint32_t n_pixels, uint32_t spp, float4 *accum, uint32_t pixel_local0) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    float4 c = accum[pixel_local0 + p];
    for (uint32_t s = 0; s < spp; ++s) {
        float4 a = acc[(size_t)p * spp + s];
        c.x += a.x; c.y += a.y; c.z += a.z;
    }
    accum[pixel_local0 + p] = c;
}

__global__ void k_finalize(const float4 *accum, uint32_t n_pixels, uint32_t n_samples, const uint32_t *per_pixel_samples, uint32_t flags,
                           float4 *out, const uint32_t *pixel_ids_dev, uint32_t pixel_begin) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    float4 c = accum[p];
    uint32_t ns = per_pixel_samples ? per_pixel_samples[p] : n_samples;
    float4 r;
    if (flags & 1u) { r = make_float4(c.x, c.y, c.z, (float)ns); }       // RT_OUT_SUM
    else { float fs = (float)ns; r = make_float4(c.x / fs, c.y / fs, c.z / fs, 1.0f); }   // main.cpp:262-263
    size_t dst = (flags & 2u) ? (size_t)(pixel_ids_dev ? pixel_ids_dev[p] : pixel_begin + p) : (size_t)p;
    out[dst] = r;
}

// ---- adaptive sampling: RenderPixel's second loop (main.cpp:245-258) ----------------------------------------
// Phase A keeps every sample colour of the min_samples pass (the reference's scratch_buffer, main.cpp:232, 241), sample-major:
// scratch[sample * stride + pixel], so that neighbouring pixels' re-reads in k_adaptive_update coalesce.
__global__ void k_resolve_scratch(const float4 *acc, uint32_t n_pixels, uint32_t spp, float4 *accum, float4 *scratch, uint32_t stride,
                                  uint32_t pixel_local0, uint32_t samp0) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    uint32_t pl = pixel_local0 + p;
    float4 c = accum[pl];
    for (uint32_t s = 0; s < spp; ++s) {
        float4 a = acc[(size_t)p * spp + s];
        scratch[(size_t)(samp0 + s) * stride + pl] = a;
        c.x += a.x; c.y += a.y; c.z += a.z;
    }
    accum[pl] = c;
}

// CalculateVariance + Color_Distance (main.cpp:179-186, 206-222) over the first `count` samples of one pixel (vals[i * stride]), same
// operation order. `sum` = the samples added left to right from 0.0f -- which is exactly what RenderPixel's running colour holds
// before the newest sample is added, so the mean's first loop need not be re-run.
RT_DEVICE float calc_variance(const float4 *vals, size_t stride, uint32_t count, float4 sum) {
    float fc = (float)count;
    float mx = sum.x / fc, my = sum.y / fc, mz = sum.z / fc;
    float variance = 0.0f;
    for (uint32_t i = 0; i < count; ++i) {
        float4 v = vals[(size_t)i * stride];
        float d = fabsf(v.x - mx) + fabsf(v.y - my) + fabsf(v.z - mz);
        variance += d * d;
    }
    variance /= (float)(count - 1u);
    return variance;
}

__global__ void k_adaptive_init(uint32_t n_pixels, const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t *act_pixel, uint32_t *act_local,
                                uint32_t *nsamples, uint32_t max_samples) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    act_pixel[p] = pixel_ids ? pixel_ids[p] : pixel_begin + p;
    act_local[p] = p;
    nsamples[p] = max_samples;
}

// K consecutive iterations of the loop at main.cpp:246-258 for every still-active pixel, from K speculatively rendered samples
// (indices samp0 .. samp0 + K - 1; slot i * K + k holds sample samp0 + k of active pixel i). Per sample, in order, exactly the
// reference's step: the new sample is stored and added, the variance of the samples BEFORE it decides (main.cpp:253 -- the newest
// sample is excluded), and a converged pixel divides by samp although samp + 1 samples were summed (main.cpp:262).
// Samples after the stopping one were rendered for nothing: they touch neither the colour nor rays_out. Rendering them K at a time
// is what makes the waves K times larger and the launch count K times smaller than one sample per iteration.
__global__ void k_adaptive_update(const float4 *acc, const uint32_t *ray_cnt, uint32_t n_active, uint32_t samp0, uint32_t K, uint32_t max_samples,
                                  const uint32_t *act_pixel, const uint32_t *act_local, float4 *accum, float4 *scratch, size_t stride, uint32_t *nsamples,
                                  uint32_t *out_pixel, uint32_t *out_local, uint32_t *n_out, unsigned long long *rays_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < n_active;
    bool keep = false;
    uint32_t pl = 0, px = 0;
    unsigned long long rays = 0;
    if (active) {
        pl = act_local[i]; px = act_pixel[i];
        float4 *sc = scratch + pl;                          // sample s of this pixel: sc[s * stride]
        float4 c = accum[pl];
        const float variance_threshold = 0.01f;
        keep = true;
        for (uint32_t k = 0; k < K && keep; ++k) {
            const uint32_t samp = samp0 + k;
            float4 a = acc[(size_t)i * K + k];
            rays += ray_cnt[(size_t)i * K + k];
            float var = calc_variance(sc, stride, samp, c);     // c: the samples before the newest, summed in order (main.cpp:242, 251)
            sc[(size_t)samp * stride] = a;
            c.x += a.x; c.y += a.y; c.z += a.z;
            if (var <= variance_threshold) { nsamples[pl] = samp; keep = false; }
            else if (!(samp + 1u < max_samples)) keep = false;
        }
        accum[pl] = c;
    }
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, o);
    if ((threadIdx.x & 31u) == 0 && rays) atomicAdd(rays_out, rays);
    uint32_t pos = warp_push(n_out, keep);
    if (keep) { out_pixel[pos] = px; out_local[pos] = pl; }
}


// RaycastHit (raytracer.cpp:20-30) for the API: position / normal / bw / vertex0 / object from a HitRec
struct ApiHit { float t; float bw[3]; uint32_t vertex0; float position[3]; float normal[3]; int32_t object; uint32_t hit; };
static_assert(sizeof(ApiHit) == 52, "rt_hit layout");

__global__ void k_hits_to_api(DevScene S, float bias, RayQueue q, const HitRec *hits, uint32_t n, ApiHit *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    HitRec h = hits[i];
    ApiHit a;
    a.t = h.t; a.hit = h.tri >= 0 ? 1u : 0u; a.object = -1; a.vertex0 = 0;
    a.bw[0] = a.bw[1] = a.bw[2] = 0.0f;
    a.position[0] = a.position[1] = a.position[2] = 0.0f;
    a.normal[0] = a.normal[1] = a.normal[2] = 0.0f;
    if (h.tri >= 0) {
        f3 org = mk3(q.o[i]), dir = mk3(q.d[i]);
        f3 ob = org + dir * bias;
        f3 pos = ob + dir * h.t;
        float4 r0 = S.tris[h.tri].r0;
        f3 gn = normalize3(mk3(r0.x, r0.y, r0.z));
        a.bw[1] = h.v; a.bw[2] = h.w; a.bw[0] = 1.0f - h.v - h.w;
        a.position[0] = pos.x; a.position[1] = pos.y; a.position[2] = pos.z;
        a.normal[0] = gn.x; a.normal[1] = gn.y; a.normal[2] = gn.z;
        a.vertex0 = S.tri_vertex0[h.tri];
        a.object = S.tri_object[h.tri];
    }
    out[i] = a;
}

__global__ void k_upload_rays(const float *rays, uint32_t n, RayQueue q) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *rp = rays + 6 * (size_t)i;
    q.o[i] = make_float4(rp[0], rp[1], rp[2], 0.0f);
    q.d[i] = make_float4(rp[3], rp[4], rp[5], __uint_as_float(i));
}

__global__ void k_queue_to_rays(RayQueue q, uint32_t n, float *rays) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 o = q.o[i], d = q.d[i];
    float *rp = rays + 6 * (size_t)i;
    rp[0] = o.x; rp[1] = o.y; rp[2] = o.z; rp[3] = d.x; rp[4] = d.y; rp[5] = d.z;
}

__global__ void k_rng_kat(uint64_t seed, uint32_t n, uint64_t *out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    PathRng r;
    rng_seed(r, seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = rng_next(r);
}

// lights >= 1 accumulate into their own per-path buffers (single writer, no atomics); fold them into the path
// accumulator in light order once every wave of the batch has finished
__global__ void k_fold_light_acc(float4 *acc, float4 *acc_extra, uint32_t n, uint32_t stride, uint32_t n_extra) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float4 a = acc[s];
    for (uint32_t l = 0; l < n_extra; ++l) {
        float4 e = acc_extra[(size_t)l * stride + s];
        a.x += e.x; a.y += e.y; a.z += e.z;
        acc_extra[(size_t)l * stride + s] = make_float4(0, 0, 0, 0);
    }
    acc[s] = a;
}

// rt_trace_rays(RT_TRACE_ANY): rays go through the SHADOW path of the wave kernel with radiance (1, 0, 0); a path
// accumulator that stayed 0 means "occluded" == TraceRay returned true
__global__ void k_rays_to_shadow_queue(const float *rays, uint32_t n, float4 *o, float4 *dir, float4 *rad, float4 *acc, uint32_t *count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = n;
    if (i >= n) return;
    const float *rp = rays + 6 * (size_t)i;
    o[i] = make_float4(rp[0], rp[1], rp[2], __uint_as_float(i));
    dir[i] = make_float4(rp[3], rp[4], rp[5], 0.0f);          // explicit directions (API rays have no light)
    rad[i] = make_float4(1.0f, 0.0f, 0.0f, -1.0f);
    acc[i] = make_float4(0, 0, 0, 0);
}
__global__ void k_occlusion_to_api(const float4 *acc, uint32_t n, ApiHit *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ApiHit a;
    memset(&a, 0, sizeof(a));
    a.hit = acc[i].x == 0.0f ? 1u : 0u;
    a.t = FLT_MAX; a.object = -1;
    out[i] = a;
}
