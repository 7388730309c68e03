// rt_groups.cuh -- SURVEY.md 8(f) #3: the reference's BuildHierarchy (bsphere.cpp:379-444) on the GPU.
//
// This builds the REFERENCE's own hierarchy over mesh groups (one leaf per OBJ group, greedy agglomeration by smallest
// parent radius, pre-order flattening) bit for bit -- the structure rt_scene_desc.spheres / sphere_group carries and whose
// leaf order defines the equal-t tie-break. The reference's host build is O(groups^3) (~3 min at 5,000 groups);
// here every merge step evaluates all pairs in parallel (one packed 64-bit atomicMin gives the reference's
// first-minimum-in-(i, j)-order choice), which turns minutes into a fraction of a second. Leaf spheres (EigenSphere: covariance,
// Jacobi, extreme points; Ritter_Iterative: 16 shrink-and-shuffle passes with the reference generator) are sequential per
// group by nature and run one thread per group. All arithmetic repeats the reference's operation order (unfused).
#pragma once
#include "rt_common.cuh"

struct GSphere { float x, y, z, r; };

RT_DEVICE f3 gs_c(const GSphere &s) { return mk3(s.x, s.y, s.z); }

struct GM33 { float e[9]; };                                     // mathlib.h:540-599, row-major
RT_DEVICE GM33 gm_identity() { GM33 m; for (int i = 0; i < 9; ++i) m.e[i] = (i % 4 == 0) ? 1.0f : 0.0f; return m; }
RT_DEVICE GM33 gm_mul(const GM33 &a, const GM33 &b) {            // mathlib.h:652-694
    GM33 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.e[i * 3 + j] = a.e[i * 3 + 0] * b.e[0 * 3 + j] + a.e[i * 3 + 1] * b.e[1 * 3 + j] + a.e[i * 3 + 2] * b.e[2 * 3 + j];
    return r;
}
RT_DEVICE GM33 gm_transpose(const GM33 &m) {
    GM33 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.e[i * 3 + j] = m.e[j * 3 + i];
    return r;
}

RT_DEVICE void g_update_sphere(GSphere &s, f3 p) {               // bsphere.cpp:14-26
    f3 pc = p - gs_c(s);
    float sq = dot3(pc, pc);
    if (sq > (s.r * s.r)) {
        float dist = sqrtf(sq);
        float nr = (float)((double)((s.r + dist) * 0.5f) + 1e-2);    // `+ 1e-2` is a double addition in the reference
        float k = (nr - s.r) / dist;
        s.r = nr;
        f3 c = gs_c(s) + pc * k;
        s.x = c.x; s.y = c.y; s.z = c.z;
    }
}

RT_DEVICE GSphere g_from_children(const GSphere &s0, const GSphere &s1) {   // bsphere.cpp:248-279
    GSphere r;
    f3 v = gs_c(s1) - gs_c(s0);
    float sq = dot3(v, v);
    float dr = s1.r - s0.r;
    if ((dr * dr) >= sq) {
        r = (s1.r >= s0.r) ? s1 : s0;
    } else {
        float dist = sqrtf(sq);
        r.r = (dist + s0.r + s1.r) * 0.5f;
        f3 c = gs_c(s0);
        if (dist > 0.001f) {
            v = mk3(v.x / dist, v.y / dist, v.z / dist);
            c = c + v * (r.r - s0.r);
        }
        r.x = c.x; r.y = c.y; r.z = c.z;
    }
    r.r *= 1.0001f;
    return r;
}

// full 16-word generator (random.h:4-42): Ritter_Iterative draws thousands of numbers per group
struct GRng { uint64_t s[16]; int p; };
RT_DEVICE void g_rng_seed(GRng &r, uint64_t seed) {
    if (seed == 0) seed = 0x5555555555555555ULL;
    r.p = 0;
    uint64_t x = seed;
    for (int i = 0; i < 16; ++i) { x ^= x >> 12; x ^= x >> 25; x ^= x >> 27; r.s[i] = x * 2685821657736338717ULL; }
}
RT_DEVICE uint64_t g_rng_next(GRng &r) {
    uint64_t s0 = r.s[r.p];
    r.p = (r.p + 1) & 15;
    uint64_t s1 = r.s[r.p];
    s1 ^= s1 << 31; s1 ^= s1 >> 11; s0 &= s0 >> 30;
    r.s[r.p] = s0 ^ s1;
    return r.s[r.p] * 1181783497276652981ULL;
}

// BoundingSphere_FromMesh (bsphere.cpp:232-246) = EigenSphere (156-195) + Ritter_Iterative (197-230); one thread per group.
// pts: scratch of one float3 per group index (the reference's calloc'd copy, shuffled in place).
__global__ void k_group_leaf_spheres(const float *positions, const uint32_t *group_first, const uint32_t *idx_positions, uint32_t n_groups,
                                     float *pts, GSphere *S) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t first = group_first[g], n = group_first[g + 1] - first;
    float *P = pts + 3 * (size_t)first;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t v = idx_positions[first + i];
        P[3 * i] = positions[3 * (size_t)v]; P[3 * i + 1] = positions[3 * (size_t)v + 1]; P[3 * i + 2] = positions[3 * (size_t)v + 2];
    }
    auto pt = [&](uint32_t i) { return mk3(P[3 * i], P[3 * i + 1], P[3 * i + 2]); };
    // CovarianceMatrix (bsphere.cpp:45-81; m(2,1) is never assigned -- sic)
    float inv = 1.0f / (float)n;
    f3 c = mk3(0, 0, 0);
    for (uint32_t i = 0; i < n; ++i) c = c + pt(i);
    c = c * inv;
    float e00 = 0, e11 = 0, e22 = 0, e01 = 0, e02 = 0, e12 = 0;
    for (uint32_t i = 0; i < n; ++i) {
        f3 p = pt(i) - c;
        e00 += p.x * p.x; e11 += p.y * p.y; e22 += p.z * p.z;
        e01 += p.x * p.y; e02 += p.x * p.z; e12 += p.y * p.z;
    }
    GM33 a;
    for (int i = 0; i < 9; ++i) a.e[i] = 0.0f;
    a.e[0] = e00 * inv; a.e[4] = e11 * inv; a.e[8] = e22 * inv;
    a.e[1] = a.e[3] = e01 * inv;
    a.e[2] = a.e[6] = e02 * inv;
    a.e[5] = e12 * inv;
    // Jacobi (bsphere.cpp:104-154) with SymSchur2 (83-102)
    GM33 v = gm_identity();
    float prevoff = 0.0f;
    for (uint32_t it = 0; it < 50; ++it) {
        uint32_t p = 0, q = 1;
        for (uint32_t i = 0; i < 3; ++i)
            for (uint32_t j = 0; j < 3; ++j)
                if (i != j && fabsf(a.e[i * 3 + j]) > fabsf(a.e[p * 3 + q])) { p = i; q = j; }
        float cs, sn;
        if (fabsf(a.e[p * 3 + q]) > 0.0001f) {
            float r = (a.e[q * 3 + q] - a.e[p * 3 + p]) / (2.0f * a.e[p * 3 + q]);
            float t;
            if (r >= 0.0f) t = 1.0f / (r + sqrtf(1.0f + r * r));
            else t = -1.0f / (-r + sqrtf(1.0f + r * r));
            cs = 1.0f / sqrtf(1.0f + t * t);
            sn = cs * t;
        } else { cs = 1.0f; sn = 0.0f; }
        GM33 J = gm_identity();
        J.e[p * 3 + p] = cs; J.e[p * 3 + q] = sn; J.e[q * 3 + p] = -sn; J.e[q * 3 + q] = cs;
        v = gm_mul(v, J);
        a = gm_mul(gm_mul(gm_transpose(J), a), J);
        float off = 0.0f;
        for (uint32_t i = 0; i < 3; ++i)
            for (uint32_t j = 0; j < 3; ++j)
                if (i != j) off += a.e[i * 3 + j] * a.e[i * 3 + j];
        if (it > 2 && off >= prevoff) break;
        prevoff = off;
    }
    // EigenSphere (bsphere.cpp:162-194)
    uint32_t max_c = 0;
    float max_e = fabsf(a.e[0]);
    if (fabsf(a.e[4]) > max_e) { max_c = 1; max_e = fabsf(a.e[4]); }
    if (fabsf(a.e[8]) > max_e) { max_c = 2; max_e = fabsf(a.e[8]); }
    f3 ev = mk3(v.e[0 * 3 + max_c], v.e[1 * 3 + max_c], v.e[2 * 3 + max_c]);
    uint32_t imin = 0, imax = 0;
    float minp = FLT_MAX, maxp = -FLT_MAX;
    for (uint32_t i = 0; i < n; ++i) {
        float proj = dot3(pt(i), ev);
        if (proj < minp) { imin = i; minp = proj; }
        if (proj > maxp) { imax = i; maxp = proj; }
    }
    GSphere s;
    {
        f3 pa = n ? pt(imin) : mk3(0, 0, 0), pb = n ? pt(imax) : mk3(0, 0, 0);
        f3 ctr = (pa + pb) * 0.5f;
        f3 d = pa - pb;
        s.x = ctr.x; s.y = ctr.y; s.z = ctr.z; s.r = sqrtf(dot3(d, d)) * 0.5f;
    }
    for (uint32_t i = 0; i < n; ++i) g_update_sphere(s, pt(i));
    // Ritter_Iterative (bsphere.cpp:197-230)
    GRng rng;
    g_rng_seed(rng, 0x201701260526ull);
    GSphere s2 = s;
    for (uint32_t k = 0; k < 16; ++k) {
        s2.r *= 0.9f;
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t remaining = n - i - 1;
            if (remaining) {
                uint32_t j = (uint32_t)g_rng_next(rng) % remaining;
                j += i + 1;
                for (int q = 0; q < 3; ++q) { float t = P[3 * i + q]; P[3 * i + q] = P[3 * j + q]; P[3 * j + q] = t; }
            }
            g_update_sphere(s2, pt(i));
        }
        if (s2.r < s.r) s = s2;
    }
    for (uint32_t i = 0; i < n; ++i) g_update_sphere(s, pt(i));
    S[g] = s;
}

// FindMergeCandidates (bsphere.cpp:281-314): min over i < j of the parent radius; ties -> the first pair in (i, j) order.
// Packed key (radius bits << 32 | i << 16 | j) orders exactly like that for positive radii and m <= 65535.
__global__ void __launch_bounds__(256) k_group_pair_min(uint32_t m, const uint32_t *list, const GSphere *S, unsigned long long *best) {
    __shared__ unsigned long long sh[8];
    uint32_t i = blockIdx.x;
    unsigned long long key = ~0ull;
    if (i < m) {
        GSphere si = S[list[i]];
        for (uint32_t j = i + 1 + threadIdx.x; j < m; j += blockDim.x) {
            GSphere p = g_from_children(si, S[list[j]]);
            unsigned long long k = ((unsigned long long)__float_as_uint(p.r) << 32) | ((unsigned long long)i << 16) | j;
            key = k < key ? k : key;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { unsigned long long other = __shfl_down_sync(0xffffffffu, key, o); key = other < key ? other : key; }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) key = sh[w] < key ? sh[w] : key;
        if (key != ~0ull) atomicMin(best, key);
    }
}

// the merge itself (bsphere.cpp:405-426): erase both, append the parent; one block
__global__ void __launch_bounds__(1024) k_group_apply_merge(uint32_t m, const uint32_t *list, uint32_t *list_out, GSphere *S, int32_t *c0, int32_t *c1,
                                                           uint32_t created, unsigned long long *best) {
    unsigned long long key = *best;
    uint32_t i = (uint32_t)((key >> 16) & 0xFFFFu), j = (uint32_t)(key & 0xFFFFu);
    uint32_t a = list[i], b = list[j];
    for (uint32_t k = threadIdx.x; k + 2 < m; k += blockDim.x) {
        uint32_t src = k;
        if (src >= i) src++;
        if (src >= j) src++;
        list_out[k] = list[src];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        S[created] = g_from_children(S[a], S[b]);
        c0[created] = (int32_t)a; c1[created] = (int32_t)b;
        list_out[m - 2] = created;
        *best = ~0ull;
    }
}
