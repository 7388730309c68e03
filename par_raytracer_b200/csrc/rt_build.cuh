// rt_build.cuh -- GPU construction of the bounding-sphere cluster hierarchy (kernel group K2).
//
// The reference builds its hierarchy on the host by greedy agglomeration: repeatedly merge the pair
// of spheres whose enclosing sphere has the smallest radius (bsphere.cpp:281-314, 399-428; O(n^3),
// leaves = whole mesh groups). This file is the B200 form of the same idea at triangle granularity:
//   1. per-triangle minimal enclosing spheres + 63-bit Morton keys of their centres,
//   2. a bitonic sort of (key, triangle) pairs,
//   3. PLOC-style parallel agglomeration: every cluster looks +-PLOC_RADIUS neighbours along the Morton
//      order for the partner with the smallest merged bounds (the reference's heuristic, bsphere.cpp:295-299,
//      measured on the merged box); mutual choices merge; survivors are compacted with a prefix sum,
//   4. subtrees of <= RT_LEAF_MAX triangles collapse into clusters; nodes are laid out in depth-first
//      pre-order (every subtree contiguous in memory) with both child bounds stored in the parent,
//   5. triangles are gathered into cluster order as SoA float4 records.
// Nothing here is bit-compared with the reference: any conservative hierarchy yields the same hits.
#pragma once
#include "rt_common.cuh"
#include "rt_sort.cuh"

#ifndef PLOC_RADIUS
#define PLOC_RADIUS 2            // search window along the Morton order. Measured (trace ms, config 2 / 1 M / 10 M triangles): radius 1: 19.2 / 15.3 / 22.9,
#endif                           // 2: 18.6 / 15.2 / 22.7, 3: 18.9 / 15.2 / 23.2, 8: 19.6 / - / 23.1, 16: 19.5 / 16.2 / 23.8, 32: 19.2 / - / 24.5; strict pairs: 41 / 44 / 71

// ---- small utilities -----------------------------------------------------------------------

RT_DEVICE uint32_t float_flip(float f) {        // order-preserving float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
RT_DEVICE float float_unflip(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

RT_DEVICE uint64_t expand21(uint32_t v) {       // spread 21 bits to every third bit
    uint64_t x = v & 0x1FFFFFu;
    x = (x | x << 32) & 0x1F00000000FFFFULL;
    x = (x | x << 16) & 0x1F0000FF0000FFULL;
    x = (x | x << 8) & 0x100F00F00F00F00FULL;
    x = (x | x << 4) & 0x10C30C30C30C30C3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

// Smallest sphere around two spheres (cf. BoundingSphere_FromChildren, bsphere.cpp:248-279), padded so
// that float rounding can never make it miss a child.
RT_DEVICE float4 enclose_spheres(float4 a, float4 b) {
    float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    float d = sqrtf(dx * dx + dy * dy + dz * dz);
    if (d + b.w <= a.w) return a;
    if (d + a.w <= b.w) return b;
    float r = 0.5f * (d + a.w + b.w);
    float k = d > 0.0f ? (r - a.w) / d : 0.0f;
    float4 o = make_float4(a.x + dx * k, a.y + dy * k, a.z + dz * k, 0.0f);
    // radius = exact worst case from the rounded centre
    float ex = o.x - a.x, ey = o.y - a.y, ez = o.z - a.z;
    float fx = o.x - b.x, fy = o.y - b.y, fz = o.z - b.z;
    float ra = sqrtf(ex * ex + ey * ey + ez * ez) + a.w;
    float rb = sqrtf(fx * fx + fy * fy + fz * fz) + b.w;
    o.w = fmaxf(ra, rb) * 1.000002f + 1e-30f;
    return o;
}
RT_DEVICE float enclose_radius(float4 a, float4 b) {
    float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    float d = sqrtf(dx * dx + dy * dy + dz * dz);
    if (d + b.w <= a.w) return a.w;
    if (d + a.w <= b.w) return b.w;
    return 0.5f * (d + a.w + b.w);
}

// ---- 1. per-triangle setup -------------------------------------------------------------------

struct BuildInput {
    const float *positions;
    const uint32_t *idx_positions;     // concatenated group index buffers
    const uint32_t *group_first;       // n_groups + 1 (index units)
    uint32_t n_groups;
    uint32_t n_tris;
};


__global__ void k_tri_spheres(BuildInput in, float4 *tri_sphere, float4 *tri_lo, float4 *tri_hi, float4 *tri_nrm, float4 *tri_slab,
                              uint32_t *bounds /*6 flipped floats*/) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    f3 lo = mk3(FLT_MAX, FLT_MAX, FLT_MAX), hi = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (j < in.n_tris) {
        f3 a = ld3(in.positions, in.idx_positions[3 * (size_t)j + 0]);
        f3 b = ld3(in.positions, in.idx_positions[3 * (size_t)j + 1]);
        f3 c = ld3(in.positions, in.idx_positions[3 * (size_t)j + 2]);
        // minimal enclosing sphere: midpoint of the longest edge if the opposite vertex is inside, else circumsphere
        f3 ab = b - a, ac = c - a, bc = c - b;
        float lab = dot3(ab, ab), lac = dot3(ac, ac), lbc = dot3(bc, bc);
        f3 p0 = a, p1 = b, p2 = c; float l = lab;
        if (lac > l) { p0 = a; p1 = c; p2 = b; l = lac; }
        if (lbc > l) { p0 = b; p1 = c; p2 = a; l = lbc; }
        f3 ctr = (p0 + p1) * 0.5f;
        f3 dv = p2 - ctr;
        if (dot3(dv, dv) > 0.25f * l) {
            // acute: circumcentre = a + (|ac|^2 (n x ab) + |ab|^2 (ac x n)) / (2 |n|^2), n = ab x ac
            f3 n = cross3(ab, ac);
            float n2 = dot3(n, n);
            if (n2 > 0.0f) {
                f3 t = cross3(n, ab) * lac + cross3(ac, n) * lab;
                ctr = a + t * (0.5f / n2);
            }
        }
        f3 da = a - ctr, db = b - ctr, dc = c - ctr;
        float r2 = fmaxf(dot3(da, da), fmaxf(dot3(db, db), dot3(dc, dc)));
        if (!(r2 < FLT_MAX)) { ctr = (a + b + c) * (1.0f / 3.0f); da = a - ctr; db = b - ctr; dc = c - ctr;
                                r2 = fmaxf(dot3(da, da), fmaxf(dot3(db, db), dot3(dc, dc))); }
        tri_sphere[j] = make_float4(ctr.x, ctr.y, ctr.z, sqrtf(r2) * 1.000002f + 1e-30f);
        {
            f3 nn = cross3(b - a, c - a);                                  // area-weighted normal (sums give a subtree's mean normal)
            tri_nrm[j] = make_float4(nn.x, nn.y, nn.z, 0.0f);
            float l = sqrtf(dot3(nn, nn));
            if (l > 0.0f && l < FLT_MAX) {
                f3 u = nn * (1.0f / l);
                float da = dot3(u, a), db = dot3(u, b), dc = dot3(u, c);
                tri_slab[j] = make_float4(u.x, u.y, u.z, fminf(da, fminf(db, dc)));
                tri_nrm[j].w = fmaxf(da, fmaxf(db, dc));               // dmax rides in the normal's w
            } else { tri_slab[j] = make_float4(0, 0, 0, -FLT_MAX); tri_nrm[j].w = FLT_MAX; }
        }
        tri_lo[j] = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.0f);
        tri_hi[j] = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.0f);
        lo = ctr; hi = ctr;
    }
    // block reduce of centre bounds
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bounds[0], float_flip(lo.x)); atomicMin(&bounds[1], float_flip(lo.y)); atomicMin(&bounds[2], float_flip(lo.z));
        atomicMax(&bounds[3], float_flip(hi.x)); atomicMax(&bounds[4], float_flip(hi.y)); atomicMax(&bounds[5], float_flip(hi.z));
    }
}

__global__ void k_morton(uint32_t n, uint32_t n_pad, const float4 *tri_sphere, const uint32_t *bounds, uint64_t *keys, uint32_t *vals) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    if (j >= n) { keys[j] = ~0ULL; vals[j] = ~0u; return; }
    float lox = float_unflip(bounds[0]), loy = float_unflip(bounds[1]), loz = float_unflip(bounds[2]);
    float hix = float_unflip(bounds[3]), hiy = float_unflip(bounds[4]), hiz = float_unflip(bounds[5]);
    float ext = fmaxf(hix - lox, fmaxf(hiy - loy, hiz - loz));
    float s = ext > 0.0f ? 2097151.0f / ext : 0.0f;
    float4 c = tri_sphere[j];
    uint32_t qx = (uint32_t)fminf(fmaxf((c.x - lox) * s, 0.0f), 2097151.0f);
    uint32_t qy = (uint32_t)fminf(fmaxf((c.y - loy) * s, 0.0f), 2097151.0f);
    uint32_t qz = (uint32_t)fminf(fmaxf((c.z - loz) * s, 0.0f), 2097151.0f);
    keys[j] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
    vals[j] = j;
}


// ---- 3. PLOC agglomeration --------------------------------------------------------------------------

struct TempTree {           // 2n - 1 nodes: [0, n) = sorted triangles, [n, 2n-1) = merges in creation order
    int32_t *c0, *c1, *parent;
    uint32_t *size;          // triangles in subtree
    uint32_t *kept;          // internal nodes with size > RT_LEAF_MAX in subtree (incl. self)
    float4 *sphere;
    float4 *lo, *hi;         // exact axis-aligned bounds of the subtree's vertices
    float4 *nsum;            // xyz: sum of area-weighted triangle normals of the subtree; w: slab dmax
    float4 *slab;            // unit normal xyz + dmin (0, 0, 0, -FLT_MAX: no slab)
};

__global__ void k_ploc_init(uint32_t n, const uint32_t *sorted_tri, const float4 *tri_sphere, const float4 *tri_lo, const float4 *tri_hi,
                            const float4 *tri_nrm, const float4 *tri_slab, int32_t *cl_node, TempTree t) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t j = sorted_tri[i];
    cl_node[i] = (int32_t)i;
    t.c0[i] = -1; t.c1[i] = -1; t.parent[i] = -1; t.size[i] = 1; t.kept[i] = 0; t.sphere[i] = tri_sphere[j];
    t.lo[i] = tri_lo[j]; t.hi[i] = tri_hi[j];
    t.nsum[i] = tri_nrm[j]; t.slab[i] = tri_slab[j];
}

// Search cost of a candidate pair = size of the merged bounds: mode 0 = squared diagonal, i.e. (2 x radius)^2 of the sphere around the
// merged box -- the reference's "smallest parent radius" criterion (bsphere.cpp:295-299) on exact extents instead of on spheres of
// spheres; mode 2 (default) = half surface area of the merged box, the right measure once the traversal bound is the box itself
// (measured: 2-4 % fewer node visits, profiles/README.md); mode 1 = strict pairing fallback for over-deep trees.
__global__ void __launch_bounds__(256) k_ploc_nn(uint32_t m, const int32_t *cl_node, TempTree t, uint32_t *nn, int pair_mode) {
    __shared__ float4 slo[256 + 2 * PLOC_RADIUS];
    __shared__ float4 shi[256 + 2 * PLOC_RADIUS];
    int base = (int)(blockIdx.x * 256) - PLOC_RADIUS;
    for (int k = threadIdx.x; k < 256 + 2 * PLOC_RADIUS; k += 256) {
        int g = base + k;
        if (g >= 0 && g < (int)m) { int32_t nd = cl_node[g]; slo[k] = t.lo[nd]; shi[k] = t.hi[nd]; }
        else { slo[k] = make_float4(0, 0, 0, -1.0f); shi[k] = slo[k]; }
    }
    __syncthreads();
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    if (pair_mode == 1) { uint32_t j = i ^ 1u; nn[i] = j < m ? j : i; return; }
    float4 lo = slo[threadIdx.x + PLOC_RADIUS], hi = shi[threadIdx.x + PLOC_RADIUS];
    float best = FLT_MAX; uint32_t bj = i;
    for (int o = -PLOC_RADIUS; o <= PLOC_RADIUS; ++o) {
        if (o == 0) continue;
        float4 l2 = slo[threadIdx.x + PLOC_RADIUS + o], h2 = shi[threadIdx.x + PLOC_RADIUS + o];
        if (l2.w < 0.0f) continue;
        float dx = fmaxf(hi.x, h2.x) - fminf(lo.x, l2.x), dy = fmaxf(hi.y, h2.y) - fminf(lo.y, l2.y), dz = fmaxf(hi.z, h2.z) - fminf(lo.z, l2.z);
        float c = pair_mode == 2 ? dx * dy + dy * dz + dz * dx : dx * dx + dy * dy + dz * dz;     // 2: half surface area of the merged box
        if (c < best) { best = c; bj = (uint32_t)((int)i + o); }
    }
    nn[i] = bj;
}

// flags packed as (merge << 32) | valid.  A mutual pair (i, nn[i]) merges into the lower position.
__global__ void k_ploc_flags(uint32_t m, const uint32_t *nn, uint64_t *flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t j = nn[i];
    bool mutual = j != i && nn[j] == i;
    uint64_t valid = (mutual && i > j) ? 0 : 1;
    uint64_t merge = (mutual && i < j) ? 1 : 0;
    flags[i] = (merge << 32) | valid;
}

__global__ void k_ploc_merge(uint32_t m, uint32_t n, uint32_t nodes_created, const uint32_t *nn, const uint64_t *flags, const uint64_t *scan,
                             const int32_t *cl_node, int32_t *out_node, TempTree t) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint64_t f = flags[i];
    if ((f & 1ull) == 0) return;
    uint64_t sc = scan[i];
    uint32_t pos = (uint32_t)(sc & 0xffffffffull);
    if (f >> 32) {
        uint32_t j = nn[i];
        int32_t a = cl_node[i], b = cl_node[j];
        int32_t id = (int32_t)(n + nodes_created + (uint32_t)(sc >> 32));
        t.c0[id] = a; t.c1[id] = b; t.parent[id] = -1; t.parent[a] = id; t.parent[b] = id;
        uint32_t sz = t.size[a] + t.size[b];
        t.size[id] = sz;
        t.kept[id] = sz > RT_LEAF_MAX ? 1u + t.kept[a] + t.kept[b] : 0u;
        t.sphere[id] = enclose_spheres(t.sphere[a], t.sphere[b]);
        float4 la = t.lo[a], lb = t.lo[b], ha = t.hi[a], hb = t.hi[b];
        t.lo[id] = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.0f);
        t.hi[id] = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.0f);
        float4 na = t.nsum[a], nb = t.nsum[b];
        t.nsum[id] = make_float4(na.x + nb.x, na.y + nb.y, na.z + nb.z, FLT_MAX);
        t.slab[id] = make_float4(0, 0, 0, -FLT_MAX);
        out_node[pos] = id;
    } else {
        out_node[pos] = cl_node[i];
    }
}

// ---- 3b. tree rotations ------------------------------------------------------------------------------
// One bottom-up sweep of local restructuring (Kensler-style rotations): at every internal node N = {L, R}, swapping L with a
// grandchild under R (or R with a grandchild under L) is applied when it shrinks the surface area of the child it rebuilds. Leaves walk
// up; the second thread to arrive at a node processes it, so everything below N is final and nobody else touches it. Only the
// rebuilt child changes bounds / size / kept; N's own bounds do not. Any tree is a valid hierarchy -- this only buys fewer node visits.
RT_DEVICE float box_half_area(float4 lo, float4 hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
RT_DEVICE float union_half_area(float4 la, float4 ha, float4 lb, float4 hb) {
    float dx = fmaxf(ha.x, hb.x) - fminf(la.x, lb.x), dy = fmaxf(ha.y, hb.y) - fminf(la.y, lb.y), dz = fmaxf(ha.z, hb.z) - fminf(la.z, lb.z);
    return dx * dy + dy * dz + dz * dx;
}
RT_DEVICE void temp_node_rebuild(TempTree &t, int32_t v) {          // bounds / size / kept / nsum of v from its two children
    const int32_t a = t.c0[v], b = t.c1[v];
    float4 la = t.lo[a], lb = t.lo[b], ha = t.hi[a], hb = t.hi[b];
    t.lo[v] = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.0f);
    t.hi[v] = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.0f);
    const uint32_t sz = t.size[a] + t.size[b];
    t.size[v] = sz;
    t.kept[v] = sz > RT_LEAF_MAX ? 1u + t.kept[a] + t.kept[b] : 0u;
    t.sphere[v] = enclose_spheres(t.sphere[a], t.sphere[b]);
    float4 na = t.nsum[a], nb = t.nsum[b];
    t.nsum[v] = make_float4(na.x + nb.x, na.y + nb.y, na.z + nb.z, FLT_MAX);
    t.slab[v] = make_float4(0, 0, 0, -FLT_MAX);
}
__global__ void k_rotate(uint32_t n, TempTree t, uint32_t *visit, uint32_t *n_rotations) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t N = t.parent[i];
    while (N >= 0) {
        __threadfence();
        if (atomicAdd(&visit[N], 1u) == 0u) return;                 // the sibling subtree is not finished: its thread will carry on
        __threadfence();
        const int32_t L = t.c0[N], R = t.c1[N];
        const bool Li = L >= (int32_t)n, Ri = R >= (int32_t)n;       // internal temp nodes live at [n, 2n - 1)
        float best = 0.0f; int which = 0;
        const float4 Llo = t.lo[L], Lhi = t.hi[L], Rlo = t.lo[R], Rhi = t.hi[R];
        if (Ri) {
            const int32_t a = t.c0[R], b = t.c1[R];
            const float now = box_half_area(Rlo, Rhi);
            const float g1 = now - union_half_area(Llo, Lhi, t.lo[b], t.hi[b]);      // L <-> a : R becomes {L, b}
            const float g2 = now - union_half_area(t.lo[a], t.hi[a], Llo, Lhi);      // L <-> b : R becomes {a, L}
            if (g1 > best) { best = g1; which = 1; }
            if (g2 > best) { best = g2; which = 2; }
        }
        if (Li) {
            const int32_t c = t.c0[L], d = t.c1[L];
            const float now = box_half_area(Llo, Lhi);
            const float g3 = now - union_half_area(Rlo, Rhi, t.lo[d], t.hi[d]);      // R <-> c : L becomes {R, d}
            const float g4 = now - union_half_area(t.lo[c], t.hi[c], Rlo, Rhi);      // R <-> d : L becomes {c, R}
            if (g3 > best) { best = g3; which = 3; }
            if (g4 > best) { best = g4; which = 4; }
        }
        if (which == 1 || which == 2) {
            const int32_t x = which == 1 ? t.c0[R] : t.c1[R];
            if (which == 1) t.c0[R] = L; else t.c1[R] = L;
            t.c0[N] = x; t.parent[x] = N; t.parent[L] = R;
            temp_node_rebuild(t, R);
            atomicAdd(n_rotations, 1u);
        } else if (which == 3 || which == 4) {
            const int32_t x = which == 3 ? t.c0[L] : t.c1[L];
            if (which == 3) t.c0[L] = R; else t.c1[L] = R;
            t.c1[N] = x; t.parent[x] = N; t.parent[R] = L;
            temp_node_rebuild(t, L);
            atomicAdd(n_rotations, 1u);
        }
        {   // N keeps its bounds and size; its kept count follows its (possibly rebuilt) children
            const uint32_t sz = t.size[N];
            t.kept[N] = sz > RT_LEAF_MAX ? 1u + t.kept[t.c0[N]] + t.kept[t.c1[N]] : 0u;
        }
        N = t.parent[N];
    }
}

// ---- 4. layout -------------------------------------------------------------------------------------

// Per temp node: first triangle slot of its subtree in cluster order, pre-order index among kept nodes, depth.
__global__ void k_layout(uint32_t n_nodes_total, TempTree t, uint32_t *tri_offset, uint32_t *kept_index, uint32_t *max_depth) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes_total) return;
    uint32_t off = 0, idx = 0, depth = 0;
    int32_t cur = (int32_t)v;
    int32_t p = t.parent[cur];
    while (p >= 0) {
        bool right = t.c1[p] == cur;
        if (right) { off += t.size[t.c0[p]]; idx += t.kept[t.c0[p]]; }
        idx += 1;
        depth++;
        cur = p; p = t.parent[cur];
    }
    tri_offset[v] = off;
    kept_index[v] = idx;
    if (t.size[v] > RT_LEAF_MAX) atomicMax(max_depth, depth + 1);
}

// cluster-order slot -> input triangle
__global__ void k_slot_to_tri(uint32_t n, const uint32_t *sorted_tri, const uint32_t *tri_offset, uint32_t *slot_tri) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_tri[tri_offset[i]] = sorted_tri[i];
}

// Tight refit: a sphere of spheres of spheres ... grows with every level. Each internal node therefore also gets the
// sphere centred on its box with the exact largest vertex distance as radius (one warp scans the node's contiguous
// triangle range); the smaller of the two survives. Both enclose every vertex of the subtree.
__global__ void __launch_bounds__(256) k_refit(uint32_t n, uint32_t n_total, TempTree t, const uint32_t *tri_offset, const uint32_t *slot_tri,
                                               BuildInput in) {
    uint32_t v = n + (blockIdx.x * blockDim.x + threadIdx.x) / 32u;
    uint32_t lane = threadIdx.x & 31u;
    if (v >= n_total) return;
    float4 lo = t.lo[v], hi = t.hi[v];
    f3 ctr = mk3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
    uint32_t off = tri_offset[v], size = t.size[v];
    float4 ns = t.nsum[v];
    float nl = sqrtf(ns.x * ns.x + ns.y * ns.y + ns.z * ns.z);
    bool has_n = nl > 0.0f && nl < FLT_MAX;
    f3 u = has_n ? mk3(ns.x / nl, ns.y / nl, ns.z / nl) : mk3(0, 0, 0);
    float d2 = 0.0f, pmin = FLT_MAX, pmax = -FLT_MAX;
    for (uint32_t k = lane; k < size; k += 32u) {
        uint32_t j = slot_tri[off + k];
        for (int c = 0; c < 3; ++c) {
            f3 q = ld3(in.positions, in.idx_positions[3 * (size_t)j + c]);
            f3 p = q - ctr;
            d2 = fmaxf(d2, dot3(p, p));
            float pr = dot3(u, q);
            pmin = fminf(pmin, pr); pmax = fmaxf(pmax, pr);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o));
        pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
        pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    }
    if (lane == 0) {
        float r = sqrtf(d2) * 1.000002f + 1e-30f;
        if (r < t.sphere[v].w) t.sphere[v] = make_float4(ctr.x, ctr.y, ctr.z, r);
        // a slab thicker than ~the sphere prunes nothing: leave it disabled
        if (has_n && (pmax - pmin) < 1.5f * t.sphere[v].w) { t.slab[v] = make_float4(u.x, u.y, u.z, pmin); t.nsum[v].w = pmax; }
    }
}

__global__ void k_emit_nodes(uint32_t n, uint32_t n_nodes_total, TempTree t, const uint32_t *tri_offset, const uint32_t *kept_index,
                             HNode *nodes, BNode *bnodes, QNode *qnodes, double qbx, double qby, double qbz, double qsx, double qsy, double qsz) {
    uint32_t v = n + blockIdx.x * blockDim.x + threadIdx.x;   // internal temp nodes only
    if (v >= n_nodes_total) return;
    if (t.size[v] <= RT_LEAF_MAX) return;
    int32_t a = t.c0[v], b = t.c1[v];
    const int32_t ref_a = t.size[a] > RT_LEAF_MAX ? (int32_t)kept_index[a] : leaf_ref(tri_offset[a], t.size[a]);
    const int32_t ref_b = t.size[b] > RT_LEAF_MAX ? (int32_t)kept_index[b] : leaf_ref(tri_offset[b], t.size[b]);
    if (bnodes || qnodes) {
        float4 la = t.lo[a], ha = t.hi[a], lb = t.lo[b], hb = t.hi[b];      // exact vertex extents; the traversal pads them per ray
        BNode q;
        // centre + half-extent per child; the half-extent is rounded UP from the rounded centre so that [c - h, c + h] contains [lo, hi]
        auto ctr = [](float lo, float hi) { return 0.5f * lo + 0.5f * hi; };
        auto hext = [](float lo, float hi, float c) { float h = fmaxf(hi - c, c - lo); return h > 0.0f ? __uint_as_float(__float_as_uint(h * 1.0000002f) + 1u) : 0.0f; };
        const float cax = ctr(la.x, ha.x), cay = ctr(la.y, ha.y), caz = ctr(la.z, ha.z), cbx = ctr(lb.x, hb.x), cby = ctr(lb.y, hb.y), cbz = ctr(lb.z, hb.z);
        q.a = make_float4(cax, cay, caz, hext(la.x, ha.x, cax)); q.b = make_float4(hext(la.y, ha.y, cay), hext(la.z, ha.z, caz), cbx, cby);
        q.c = make_float4(cbz, hext(lb.x, hb.x, cbx), hext(lb.y, hb.y, cby), hext(lb.z, hb.z, cbz));
        q.c0 = ref_a; q.c1 = ref_b;
        q.pad0 = q.pad1 = 0;
        if (bnodes) bnodes[kept_index[v]] = q;
        // 15-bit grid: lo -> floor, hi -> ceil, in double so that the grid plane is never inside the exact extent
        QNode z;
        auto ql = [](float x, double b, double st) { double g = floor(((double)x - b) / st); return (uint32_t)fmin(fmax(g, 0.0), 32767.0); };
        auto qh = [](float x, double b, double st) { double g = ceil(((double)x - b) / st); return (uint32_t)fmin(fmax(g, 0.0), 32767.0); };
        z.w[0] = ql(la.x, qbx, qsx) | (qh(ha.x, qbx, qsx) << 16); z.w[1] = ql(la.y, qby, qsy) | (qh(ha.y, qby, qsy) << 16);
        z.w[2] = ql(la.z, qbz, qsz) | (qh(ha.z, qbz, qsz) << 16); z.w[3] = (uint32_t)q.c0;
        z.w[4] = ql(lb.x, qbx, qsx) | (qh(hb.x, qbx, qsx) << 16); z.w[5] = ql(lb.y, qby, qsy) | (qh(hb.y, qby, qsy) << 16);
        z.w[6] = ql(lb.z, qbz, qsz) | (qh(hb.z, qbz, qsz) << 16); z.w[7] = (uint32_t)q.c1;
        if (qnodes) qnodes[kept_index[v]] = z;
    }
    if (!nodes) return;
    HNode o;
    o.s0 = t.sphere[a]; o.s1 = t.sphere[b];
    o.p0 = t.slab[a]; o.p1 = t.slab[b];
    o.dmax0 = t.nsum[a].w; o.dmax1 = t.nsum[b].w;
    o.c0 = ref_a; o.c1 = ref_b;
    nodes[kept_index[v]] = o;
}

// ---- 4b. collapse to the 4-wide form -------------------------------------------------------------------------
// Level-synchronous, top-down: the frontier holds the binary nodes that become wide nodes of this level. k_wide_expand fills a wide
// node's four slots -- start from the two binary children, replace the internal child with the largest box by ITS two children until
// four slots are used or only clusters are left -- and counts the internal children; a prefix sum over the frontier gives every
// internal child its index in the next level (children of one node contiguous, levels one after the other: neighbours in a level are
// neighbours in space, because the order is inherited from the Morton order of the parents).
struct WideTree {
    int32_t *src;        // [n_wide] binary node id of every wide node
    int32_t *child;      // [4 * n_wide] binary node ids of the slots (-1: empty)
    int32_t *ref;        // [4 * n_wide] child refs as the traversal reads them
};

__global__ void k_wide_expand(uint32_t m, uint32_t first, TempTree t, WideTree w, uint64_t *counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t X = w.src[first + i];
    int32_t ch[4] = {t.c0[X], t.c1[X], -1, -1};
    int n = 2;
    while (n < 4) {
        int best = -1; float best_area = -1.0f;
        for (int k = 0; k < n; ++k) {
            if (t.size[ch[k]] <= RT_LEAF_MAX) continue;
            float a = box_half_area(t.lo[ch[k]], t.hi[ch[k]]);
            if (a > best_area) { best_area = a; best = k; }
        }
        if (best < 0) break;
        const int32_t e = ch[best];
        ch[best] = t.c0[e]; ch[n++] = t.c1[e];
    }
    uint32_t internal = 0;
    for (int k = 0; k < 4; ++k) {
        w.child[4 * (size_t)(first + i) + k] = k < n ? ch[k] : -1;
        if (k < n && t.size[ch[k]] > RT_LEAF_MAX) internal++;
    }
    counts[i] = internal;
}

__global__ void k_wide_assign(uint32_t m, uint32_t first, uint32_t next_first, TempTree t, WideTree w, const uint64_t *scan, const uint32_t *tri_offset) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t at = next_first + (uint32_t)scan[i];
    for (int k = 0; k < 4; ++k) {
        const int32_t c = w.child[4 * (size_t)(first + i) + k];
        int32_t ref = (int32_t)RT_EMPTY_REF;
        if (c >= 0) {
            if (t.size[c] > RT_LEAF_MAX) { w.src[at] = c; ref = (int32_t)at; at++; }
            else ref = leaf_ref(tri_offset[c], t.size[c]);
        }
        w.ref[4 * (size_t)(first + i) + k] = ref;
    }
}

RT_DEVICE uint32_t quant_lo(float x, double b, double st) { double g = floor(((double)x - b) / st); return (uint32_t)fmin(fmax(g, 0.0), 32767.0); }
RT_DEVICE uint32_t quant_hi(float x, double b, double st) { double g = ceil(((double)x - b) / st); return (uint32_t)fmin(fmax(g, 0.0), 32767.0); }

__global__ void k_wide_emit(uint32_t n_wide, TempTree t, WideTree w, Q4Node *out, double qbx, double qby, double qbz, double qsx, double qsy, double qsz) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_wide) return;
    Q4Node z;
    for (int k = 0; k < 4; ++k) {
        const int32_t c = w.child[4 * (size_t)i + k];
        if (c < 0) { z.c[k] = make_uint4(32767u, 32767u, 32767u, RT_EMPTY_REF); continue; }     // lo = 32767, hi = 0: an inverted box (and the ref says empty)
        const float4 lo = t.lo[c], hi = t.hi[c];                   // exact vertex extents; lo -> floor, hi -> ceil, in double (as k_emit_nodes)
        z.c[k] = make_uint4(quant_lo(lo.x, qbx, qsx) | (quant_hi(hi.x, qbx, qsx) << 16), quant_lo(lo.y, qby, qsy) | (quant_hi(hi.y, qby, qsy) << 16),
                            quant_lo(lo.z, qbz, qsz) | (quant_hi(hi.z, qbz, qsz) << 16), (uint32_t)w.ref[4 * (size_t)i + k]);
    }
    out[i] = z;
}

// ---- 5. gather triangles into cluster order ----------------------------------------------------------

struct GatherInput {
    const float *positions, *texcoords, *normals, *tangents;
    const uint32_t *idx_p, *idx_t, *idx_n;
    const uint32_t *group_first;
    const uint32_t *group_rank_base;    // rank of the group's first triangle in the reference's visit order
    const int32_t *group_object;        // sphere index holding the group
    const int32_t *group_material;      // resolved (default -> n_materials)
    uint32_t n_groups, n_tris;
};


__global__ void k_gather(GatherInput in, const uint32_t *sorted_tri, const uint32_t *tri_offset, TriRec *tris, uint32_t *tri_rank,
                         float4 *tri_uv, float4 *tri_nrm, float4 *tri_tan, uint32_t *tri_vertex0, int32_t *tri_object, uint8_t *tri_mat) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;      // sorted position == temp leaf node id
    if (i >= in.n_tris) return;
    uint32_t j = sorted_tri[i];                               // input triangle
    uint32_t dst = tri_offset[i];
    uint32_t g = find_group(in.group_first, in.n_groups, 3u * j);
    uint32_t v0 = 3u * j - in.group_first[g];
    f3 a = ld3(in.positions, in.idx_p[3 * (size_t)j + 0]);
    f3 b = ld3(in.positions, in.idx_p[3 * (size_t)j + 1]);
    f3 c = ld3(in.positions, in.idx_p[3 * (size_t)j + 2]);
    f3 ab = b - a;                                            // raytracer.cpp:85-86, 91 -- unfused, exact
    f3 ac = c - a;
    f3 n = cross3(ab, ac);
    TriRec r;
    r.r0 = make_float4(n.x, n.y, n.z, a.x);
    r.r1 = make_float4(a.y, a.z, ab.x, ab.y);
    r.r2 = make_float4(ab.z, ac.x, ac.y, ac.z);
    tris[dst] = r;
    tri_rank[dst] = in.group_rank_base[g] + v0 / 3u;
    tri_vertex0[dst] = v0;
    tri_object[dst] = in.group_object[g];
    tri_mat[dst] = (uint8_t)((uint32_t)in.group_material[g] & 255u);
    const float *t0 = in.texcoords + 2 * (size_t)in.idx_t[3 * (size_t)j + 0];
    const float *t1 = in.texcoords + 2 * (size_t)in.idx_t[3 * (size_t)j + 1];
    const float *t2 = in.texcoords + 2 * (size_t)in.idx_t[3 * (size_t)j + 2];
    tri_uv[2 * (size_t)dst + 0] = make_float4(t0[0], t0[1], t1[0], t1[1]);
    tri_uv[2 * (size_t)dst + 1] = make_float4(t2[0], t2[1], __int_as_float(in.group_material[g]), 0.0f);
    f3 gn = normalize3(n);                                    // RaycastHit::normal (raytracer.cpp:122): a per-triangle constant
    const float gnk[3] = {gn.x, gn.y, gn.z};
    for (int k = 0; k < 3; ++k) {
        uint32_t ni = in.idx_n[3 * (size_t)j + k];
        f3 nn = ld3(in.normals, ni);
        tri_nrm[3 * (size_t)dst + k] = make_float4(nn.x, nn.y, nn.z, gnk[k]);
        if (tri_tan) {
            f3 tt = ld3(in.tangents, ni);
            tri_tan[3 * (size_t)dst + k] = make_float4(tt.x, tt.y, tt.z, 0.0f);
        }
    }
}
