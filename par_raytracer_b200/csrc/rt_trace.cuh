// rt_trace.cuh -- kernel group K3 / K3s: hierarchy traversal + child-bound test + ray/triangle.
//
// Replaces TraceRay (raytracer.cpp:159-232) with its callees IntersectRaySphere (32-60),
// IntersectRayMesh (127-157) and IntersectRayTriangle (82-125).
//
// What must be bit-exact and what need not be:
//  * The TRIANGLE test decides t, the barycentrics and hit/miss; it repeats the reference's float
//    operations in the reference's order (unfused; this TU is built with -fmad=false).
//  * The closest hit is the lexicographic minimum of (t, rank): rank = position of the triangle in the
//    reference's own encounter order (c1-subtree-first DFS over its hierarchy, then index order inside a
//    group; raytracer.cpp:136, 149, 208-209, 220), so equal-t ties resolve as in the reference although
//    our traversal order is different.
//  * The CHILD-BOUND test only prunes (the reference's IntersectRaySphere plays that role there). Three
//    interchangeable forms -- boxes on a 15-bit grid (default), float boxes, sphere + slab -- all fattened
//    by a slack proportional to the distance from the ray origin, so float rounding can never prune a
//    triangle the exact test would accept; FMA and approximate reciprocals are allowed in them.
#pragma once
#include "rt_common.cuh"
#include "rt_raygen.cuh"
#include "rt_types.cuh"

#define RT_CULL_SLACK 4e-6f


RT_DEVICE float approx_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RT_DEVICE float approx_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // culling only: 1 MUFU

// Conservative test of a ray against one child bound = (sphere fattened by `slack`) INTERSECT (slab fattened by `slack`).
//   sphere: distance from the centre to the ray line through the perpendicular foot (robust; no b*b - c cancellation),
//           giving the parameter interval [tca - half, tca + half], clipped to [0, tmax];
//   slab:   over that clipped interval [T0, T1], n.(o + t d) - n.o sweeps [min(T0 nd, T1 nd), max(T0 nd, T1 nd)], which must
//           meet [dmin - slack - n.o, dmax + slack - n.o]. No division, no special case for rays parallel to the slab.
// `slack` is per ray: RT_CULL_SLACK * (|o|_1 + scene bound) >= RT_CULL_SLACK * (|c - o|_1 + r) for every node, i.e. proportional
// to the largest distance at which this ray's exact triangle arithmetic can still round differently. FMA and approximate
// sqrt are allowed here: the result only prunes. Returns the sphere entry distance for near/far ordering.
RT_DEVICE bool cull_child(float4 s, float4 p, float dmax, f3 o, f3 d, float inv_dd, float slack, float tmax, float &t_entry) {
    float mx = s.x - o.x, my = s.y - o.y, mz = s.z - o.z;
    float b = __fmaf_rn(mz, d.z, __fmaf_rn(my, d.y, mx * d.x));
    float tca = b * inv_dd;
    float px = __fmaf_rn(-tca, d.x, mx), py = __fmaf_rn(-tca, d.y, my), pz = __fmaf_rn(-tca, d.z, mz);
    float dist2 = __fmaf_rn(pz, pz, __fmaf_rn(py, py, px * px));
    float r = s.w + slack;
    float h2 = __fmaf_rn(r, r, -dist2);
    float half = approx_sqrt(fmaxf(h2, 0.0f) * inv_dd);
    float t0 = tca - half, t1 = tca + half;
    t_entry = t0;
    float T0 = fmaxf(t0, 0.0f), T1 = fminf(t1, tmax);
    float no = __fmaf_rn(p.z, o.z, __fmaf_rn(p.y, o.y, p.x * o.x));
    float nd = __fmaf_rn(p.z, d.z, __fmaf_rn(p.y, d.y, p.x * d.x));
    float e0 = T0 * nd, e1 = T1 * nd;
    float lo = (p.w - slack) - no, hi = (dmax + slack) - no;
    // h2 >= 0 is false for NaN radii (empty child); T0 <= T1 <=> the sphere interval meets [0, tmax] (strict > on tmax as raytracer.cpp:177)
    return (h2 >= 0.0f) & (T0 <= T1) & !((fmaxf(e0, e1) < lo) | (fminf(e0, e1) > hi));
}

struct RayCtx {
    f3 o;      // biased origin (raytracer.cpp:163)
    f3 d;
    f3 qp;     // o - (o + d)   (raytracer.cpp:88-89: NOT -d in floats)
};

// IntersectRayTriangle (raytracer.cpp:82-125) on a precomputed record. Returns true when the triangle is
// hit at all; t/v/w are the reference's out_hit->t, bw.y, bw.z.
RT_DEVICE bool tri_test(const TriRec &r, const RayCtx &c, float &t_out, float &v_out, float &w_out) {
    f3 n = mk3(r.r0.x, r.r0.y, r.r0.z);
    float dd = dot3(c.qp, n);
    if (dd <= 0.0f) return false;
    f3 a = mk3(r.r0.w, r.r1.x, r.r1.y);
    f3 ap = c.o - a;
    float t = dot3(ap, n);
    if (t < 0.0f) return false;
    f3 e = cross3(c.qp, ap);
    f3 ac = mk3(r.r2.y, r.r2.z, r.r2.w);
    float v = dot3(ac, e);
    if (v < 0.0f || v > dd) return false;
    f3 ab = mk3(r.r1.z, r.r1.w, r.r2.x);
    float w = -dot3(ab, e);
    if (w < 0.0f || (v + w) > dd) return false;
    float ood = 1.0f / dd;
    t_out = t * ood;
    v_out = v * ood;
    w_out = w * ood;
    return true;
}

// ---- axis-aligned child bounds (default) ---------------------------------------------------------------------
// Slab test of a ray against one child box fattened by the per-ray `slack` on every side (same slack, same argument as
// cull_child). With i = 1/d per axis, the parameter of the fattened lo plane is (lo - slack - o) i = lo i - (o + slack) i and of the
// fattened hi plane (hi + slack - o) i = hi i - (o - slack) i, whatever the sign of d: one FMA per plane against two per-ray
// constants per axis, so the padding is free. |d| is clamped away from zero (1e-20: the fake drift over any t is far below
// the slack) so that no 0 * inf appears. FMA / approximate reciprocal are allowed: the result only prunes.
struct BoxRay { float ix, iy, iz, ax, ay, az, cnx, cny, cnz, cfx, cfy, cfz; };

RT_DEVICE float safe_rcp(float d) { return approx_rcp(fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d); }

// Centre / half-extent form of the float box: the parameter interval of the fattened slab of axis a is
//     [ (c - o) i - (h + slack) |i| ,  (c - o) i + (h + slack) |i| ],   i = 1 / d,
// whatever the sign of d -- no near / far selection, no per-axis min / max: four FMAs per axis against per-ray constants
// (cn = -o i - slack |i|, cf = -o i + slack |i|), all on the FMA pipe. The ALU pipe (min / max, compares, byte permutes), which bounds
// the quantised form, only sees the two 3-input min / max pairs and the compare.
RT_DEVICE void box_ray_setup(BoxRay &R, f3 o, f3 d, float slack) {
    R.ix = safe_rcp(d.x); R.iy = safe_rcp(d.y); R.iz = safe_rcp(d.z);
    R.ax = fabsf(R.ix); R.ay = fabsf(R.iy); R.az = fabsf(R.iz);
    R.cnx = __fmaf_rn(-slack, R.ax, -o.x * R.ix); R.cfx = __fmaf_rn(slack, R.ax, -o.x * R.ix);
    R.cny = __fmaf_rn(-slack, R.ay, -o.y * R.iy); R.cfy = __fmaf_rn(slack, R.ay, -o.y * R.iy);
    R.cnz = __fmaf_rn(-slack, R.az, -o.z * R.iz); R.cfz = __fmaf_rn(slack, R.az, -o.z * R.iz);
}

RT_DEVICE bool box_child(float cx, float cy, float cz, float hx, float hy, float hz, const BoxRay &R, float tmax, float &tn) {
    float tnx = __fmaf_rn(-hx, R.ax, __fmaf_rn(cx, R.ix, R.cnx)), tfx = __fmaf_rn(hx, R.ax, __fmaf_rn(cx, R.ix, R.cfx));
    float tny = __fmaf_rn(-hy, R.ay, __fmaf_rn(cy, R.iy, R.cny)), tfy = __fmaf_rn(hy, R.ay, __fmaf_rn(cy, R.iy, R.cfy));
    float tnz = __fmaf_rn(-hz, R.az, __fmaf_rn(cz, R.iz, R.cnz)), tfz = __fmaf_rn(hz, R.az, __fmaf_rn(cz, R.iz, R.cfz));
    tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));       // FMNMX3 pairs on sm_100a
    float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    return tn <= tf;
}

// ---- quantised child bounds (default) ------------------------------------------------------------------------
// Same slab test on QNode's 15-bit grid. A grid coordinate q becomes the float 0.5 + q 2^-16 with ONE byte permute
// (bytes 0x3F, q.hi, q.lo, 0x00), and plane parameter t = (qmid + 32768 step + q step - o) / d = qf * A + C with per-ray
// A = 65536 step / d and C = (qmid - o) / d: one PRMT + one FMA per plane. The permute selector picks the lo or the hi half of the
// axis word, so the ray's direction signs choose near and far planes up front and the per-axis min / max disappears: a node
// visit costs about what the float-box visit costs, with half the bytes and half the load instructions. The per-ray slack
// moves the near plane towards the origin side and the far plane away from it (cn / cf), as in box_ray_setup.
struct QRay { float ax, ay, az, cnx, cny, cnz, cfx, cfy, cfz; uint32_t snx, sny, snz; };   // s?: near selector; far = near ^ 0x0220

RT_DEVICE void qbox_ray_setup(QRay &Q, const DevScene &S, f3 o, f3 d, float slack) {
    float dx = fabsf(d.x) < 1e-20f ? copysignf(1e-20f, d.x) : d.x;
    float dy = fabsf(d.y) < 1e-20f ? copysignf(1e-20f, d.y) : d.y;
    float dz = fabsf(d.z) < 1e-20f ? copysignf(1e-20f, d.z) : d.z;
    float ix = approx_rcp(dx), iy = approx_rcp(dy), iz = approx_rcp(dz);
    bool px = !signbit(dx), py = !signbit(dy), pz = !signbit(dz);
    Q.ax = 65536.0f * S.qstep[0] * ix; Q.ay = 65536.0f * S.qstep[1] * iy; Q.az = 65536.0f * S.qstep[2] * iz;
    float sx = px ? slack : -slack, sy = py ? slack : -slack, sz = pz ? slack : -slack;
    Q.cnx = (S.qmid[0] - (o.x + sx)) * ix; Q.cfx = (S.qmid[0] - (o.x - sx)) * ix;
    Q.cny = (S.qmid[1] - (o.y + sy)) * iy; Q.cfy = (S.qmid[1] - (o.y - sy)) * iy;
    Q.cnz = (S.qmid[2] - (o.z + sz)) * iz; Q.cfz = (S.qmid[2] - (o.z - sz)) * iz;
    Q.snx = px ? 0x7104u : 0x7324u; Q.sny = py ? 0x7104u : 0x7324u; Q.snz = pz ? 0x7104u : 0x7324u;
}

RT_DEVICE bool qbox_child(uint32_t wx, uint32_t wy, uint32_t wz, const QRay &Q, float tmax, float &tn) {
    const uint32_t K = 0x3F000000u;
    float tnx = __fmaf_rn(__uint_as_float(__byte_perm(wx, K, Q.snx)), Q.ax, Q.cnx);
    float tny = __fmaf_rn(__uint_as_float(__byte_perm(wy, K, Q.sny)), Q.ay, Q.cny);
    float tnz = __fmaf_rn(__uint_as_float(__byte_perm(wz, K, Q.snz)), Q.az, Q.cnz);
    float tfx = __fmaf_rn(__uint_as_float(__byte_perm(wx, K, Q.snx ^ 0x0220u)), Q.ax, Q.cfx);
    float tfy = __fmaf_rn(__uint_as_float(__byte_perm(wy, K, Q.sny ^ 0x0220u)), Q.ay, Q.cfy);
    float tfz = __fmaf_rn(__uint_as_float(__byte_perm(wz, K, Q.snz ^ 0x0220u)), Q.az, Q.cfz);
    tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
    float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    return tn <= tf;
}

// ---- the wave trace kernel -----------------------------------------------------------------------------
// One launch traces everything a wave has to trace: the closest-hit rays of the pending recursion nodes AND the
// shadow rays queued by the previous shading step (ShadeLight, raytracer.cpp:378-411), so a wave pays one
// launch tail instead of two.
//
// Warp-cooperative scheduling: warps are persistent (grid = resident warps of the chip) and pull rays from one
// global work counter. A lane whose ray has finished does not wait for the slowest lane of its warp: as soon as
// RT_FETCH_MIN lanes are idle the warp fetches that many consecutive rays with ONE atomicAdd (consecutive
// queue entries are spatially coherent: same pixel / neighbouring pixels). Traversal per lane is a short-stack
// while-while loop: descend internal nodes nearest-child first, then scan the reached cluster. stack[0] holds a
// sentinel, so "pop" needs no emptiness test: popping the sentinel ends the ray.
//
// BOUNDS selects the child bound: axis-aligned boxes on the 15-bit scene grid (32-byte nodes, 2 loads per visit; default),
// float boxes (64-byte nodes, 4 loads, ~60 instructions per visit) or the sphere + slab bound (80-byte nodes, 5 loads, ~105
// instructions per visit). All are conservative, so hits are identical (tests run all three).
#define RT_TRACE_BLOCK 128
#ifndef RT_TRACE_MIN_BLOCKS
#define RT_TRACE_MIN_BLOCKS 8          // 64 registers per thread: 32 warps per SM
#endif
#ifndef RT_FETCH_MIN
#define RT_FETCH_MIN 16
#endif
#ifndef RT_FETCH_PRIMARY
#define RT_FETCH_PRIMARY 24         // same threshold while the warp still draws primary rays of wave 0: coherent rays finish together, so waiting for more idle
#endif                              // lanes costs little and saves refills (measured 8 / 16 / 24 / 32 on config 3: 166.6 / 163.2 / 160.9 / 161.8 ms per frame; config 4: 16 / 24 / 28: 1,557 / 1,524 / 1,518)
#ifndef RT_LEAF_WAIT
#define RT_LEAF_WAIT 12
#endif
#ifndef RT_NODE_UNROLL
#define RT_NODE_UNROLL 2            // node visits between two warp votes (measured: 2 and 4 are equal, 1 is 3-5 % slower)
#endif
#ifndef RT_STACK_TOP_REG
#define RT_STACK_TOP_REG 1          // 1: the logical stack top lives in a register (a pop is a move, the reload is issued at once)
#endif
#define RT_DONE ((int)0x80000000)      // never a leaf ref: |leaf ref| <= 1 + 8 * 2e8 + 7 < 2^31


template <bool COUNT, int BOUNDS>
__global__ void __launch_bounds__(RT_TRACE_BLOCK, RT_TRACE_MIN_BLOCKS) k_trace_wave(DevScene S, float bias, WaveQueues W, PrimaryGen G, TraceCounters *counters) {
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t nC = G.enabled ? G.n_slots : (W.n_closest ? min(*W.n_closest, W.closest_max) : W.closest_max);
    uint32_t total = nC;
    for (uint32_t l = 0; l < W.n_lights; ++l) total += min(W.n_shadow[l], W.shadow_stride);
    unsigned long long n_sph = 0, n_clu = 0;

    bool live = false, exhausted = false;
    uint32_t fetch_min = G.enabled ? W.fetch_min_primary : W.fetch_min;
    RayCtx c; c.o = mk3(0, 0, 0); c.d = c.o; c.qp = c.o;
    constexpr bool BOX = BOUNDS == RT_BOUNDS_BOX, Q4 = BOUNDS == RT_BOUNDS_QBOX4, QBOX = BOUNDS == RT_BOUNDS_QBOX || Q4;
    BoxRay R; R.ix = R.iy = R.iz = R.ax = R.ay = R.az = R.cnx = R.cny = R.cnz = R.cfx = R.cfy = R.cfz = 0.0f;
    QRay Q; Q.ax = Q.ay = Q.az = Q.cnx = Q.cny = Q.cnz = Q.cfx = Q.cfy = Q.cfz = 0.0f; Q.snx = Q.sny = Q.snz = 0x7104u;
    float inv_dd = 0.0f, slack = 0.0f;
    HitRec best; best.t = FLT_MAX; best.v = 0; best.w = 0; best.tri = -1;
    float tcull = FLT_MAX;         // boxes: best.t widened by 1e-5 relative, so a box entered a few ulps beyond best.t is still opened
    uint32_t best_rank = 0xFFFFFFFFu, out_idx = 0;
    int kind = 0;                  // 0 closest hit -> hits[out_idx]; 1 shadow, boolean; 2 shadow, needs t (point light); shadow kinds carry the light index << 2
    // Traversal stack: the logical top lives in the register `top`, the rest in local memory (stack[0] = sentinel). A pop takes
    // the register and issues the reload of the next entry at once, so the load latency is off the critical path of the descent.
    int cur = RT_DONE, top = RT_DONE;
    int stack[Q4 ? RT_STACK4_MAX : RT_STACK_MAX];
    stack[0] = RT_DONE;
    int *sptr = stack + 1;         // next free entry

    while (true) {
        // ---- warp-cooperative fetch ----
        uint32_t idle = __ballot_sync(FULL, !live);
        if (!exhausted && (idle == FULL || __popc(idle) >= fetch_min)) {
            uint32_t n_idle = __popc(idle), base = 0;
            if (lane == 0) base = atomicAdd(W.next, n_idle);
            base = __shfl_sync(FULL, base, 0);
            if (base + n_idle >= total) exhausted = true;
            if (base + n_idle >= nC) fetch_min = W.fetch_min_shadow;      // warp-uniform: the counter has moved on to shadow rays
            if (!live) {
                uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < total) {
                    float4 o4, d4;
                    if (idx < nC) {
                        kind = 0; out_idx = idx;
                        if (G.enabled) {
                            PathRng pr; f3 po, pd; primary_ray(G, idx, pr, po, pd); o4 = mk4(po, 0.0f); d4 = mk4(pd, 0.0f);
                            W.closest.d[idx] = d4;                       // wave 0: the shading step reads the direction back
                        }
                        else { o4 = W.closest.o[idx]; d4 = W.closest.d[idx]; }
                    }
                    else {
                        uint32_t j = idx - nC, light = 0;
                        while (true) { uint32_t ns = min(W.n_shadow[light], W.shadow_stride); if (j < ns) break; j -= ns; light++; }
                        size_t e = (size_t)light * W.shadow_stride + j;
                        o4 = W.shadow_o[e];
                        kind = (W.rad[e].w < 0.0f ? 1 : 2) | (int)(light << 2); out_idx = (uint32_t)e;
                        if (W.shadow_dir) d4 = W.shadow_dir[e];
                        else {
                            const DevLight &Lt = S.lights[light];
                            f3 lv = Lt.type == 0 ? mk3(Lt.facing[0], Lt.facing[1], Lt.facing[2]) * -1.0f                       // raytracer.cpp:240
                                                 : normalize3(mk3(Lt.position[0], Lt.position[1], Lt.position[2]) - mk3(o4));    // raytracer.cpp:243
                            d4 = mk4(lv, 0.0f);
                        }
                    }
                    f3 dir = mk3(d4);
                    c.o = mk3(o4) + dir * bias;                      // raytracer.cpp:163
                    f3 q = c.o + dir;
                    c.qp = c.o - q;
                    slack = RT_CULL_SLACK * (fabsf(c.o.x) + fabsf(c.o.y) + fabsf(c.o.z) + S.cull_bound);
                    if (QBOX) qbox_ray_setup(Q, S, c.o, dir, slack);
                    else if (BOX) box_ray_setup(R, c.o, dir, slack);
                    else { c.d = dir; inv_dd = approx_rcp(__fmaf_rn(dir.z, dir.z, __fmaf_rn(dir.y, dir.y, dir.x * dir.x))); }   // culling only
                    best.t = FLT_MAX; best.v = 0.0f; best.w = 0.0f; best.tri = -1; best_rank = 0xFFFFFFFFu;   // raytracer.cpp:166
                    tcull = FLT_MAX;
                    cur = S.n_tris ? S.root : RT_DONE; sptr = stack + 1; top = RT_DONE;
                    live = true;
                }
            }
        }
        if (__ballot_sync(FULL, live) == 0) break;

        // ---- descend to the next cluster ----
        // The node loop is warp-uniform: it runs while enough lanes still have an internal node to open. A lane that has reached a
        // cluster (or finished) waits, but only until `leaf_wait` lanes are waiting -- then the warp leaves the loop, scans the
        // reached clusters, refills finished lanes and comes back. Without the cap the slowest descent of the warp holds every
        // other lane (ncu: 10 of 32 lanes active in the node loop); idle lanes hold cur = RT_DONE.
        {
            const int keep = max(1, __popc(__ballot_sync(FULL, live)) - (int)W.leaf_wait);
            uint32_t nm = __ballot_sync(FULL, cur >= 0);
            while (nm != 0) {
#pragma unroll
                for (int u = 0; u < RT_NODE_UNROLL; ++u)             // node visits per warp vote
                if (Q4) {
                  if (cur >= 0) {
                    // four children: all four 16-byte words are requested together; the hit children are ordered by entry distance with a
                    // 5-exchange network on keys = (entry distance bits & ~3) | slot (distances are >= 0: their bits order like unsigned
                    // integers; a miss is the largest key), nearest first, the others pushed far to near
                    const uint4 *np = reinterpret_cast<const uint4 *>(S.q4nodes + cur);
                    const uint4 A = __ldg(np), B = __ldg(np + 1), C = __ldg(np + 2), D = __ldg(np + 3);
                    float t0, t1, t2, t3;
                    const bool h0 = qbox_child(A.x, A.y, A.z, Q, tcull, t0) & (A.w != RT_EMPTY_REF);
                    const bool h1 = qbox_child(B.x, B.y, B.z, Q, tcull, t1) & (B.w != RT_EMPTY_REF);
                    const bool h2 = qbox_child(C.x, C.y, C.z, Q, tcull, t2) & (C.w != RT_EMPTY_REF);
                    const bool h3 = qbox_child(D.x, D.y, D.z, Q, tcull, t3) & (D.w != RT_EMPTY_REF);
                    if (COUNT) n_sph += 4;
                    const uint32_t k0 = h0 ? (__float_as_uint(t0) & ~3u) : 0xFFFFFFFFu, k1 = h1 ? ((__float_as_uint(t1) & ~3u) | 1u) : 0xFFFFFFFFu;
                    const uint32_t k2 = h2 ? ((__float_as_uint(t2) & ~3u) | 2u) : 0xFFFFFFFFu, k3 = h3 ? ((__float_as_uint(t3) & ~3u) | 3u) : 0xFFFFFFFFu;
                    const uint32_t a = min(k0, k1), b = max(k0, k1), c2 = min(k2, k3), d2 = max(k2, k3);
                    const uint32_t s0 = min(a, c2), f = max(a, c2), g = min(b, d2), s3 = max(b, d2);
                    const uint32_t s1 = min(f, g), s2 = max(f, g);
                    auto pick = [&](uint32_t key) { return (int)((key & 2u) ? ((key & 1u) ? D.w : C.w) : ((key & 1u) ? B.w : A.w)); };
                    if (s0 != 0xFFFFFFFFu) {
                        if (s3 != 0xFFFFFFFFu) { *sptr++ = top; top = pick(s3); }     // 3 pushes per level at most: the build bounds the depth
                        if (s2 != 0xFFFFFFFFu) { *sptr++ = top; top = pick(s2); }
                        if (s1 != 0xFFFFFFFFu) { *sptr++ = top; top = pick(s1); }
                        cur = pick(s0);
                    } else { cur = top; top = *--sptr; }
                  }
                } else
                if (cur >= 0) {
                    int2 ch; bool h0, h1; float t0, t1;
                    if (QBOX) {
                        const uint4 *np = reinterpret_cast<const uint4 *>(S.qnodes + cur);
                        uint4 A = __ldg(np), B = __ldg(np + 1);
                        ch = make_int2((int)A.w, (int)B.w);
                        h0 = qbox_child(A.x, A.y, A.z, Q, tcull, t0);
                        h1 = qbox_child(B.x, B.y, B.z, Q, tcull, t1);
                    } else if (BOX) {
                        const float4 *np = reinterpret_cast<const float4 *>(S.bnodes + cur);
                        float4 A = __ldg(np), B = __ldg(np + 1), C = __ldg(np + 2);
                        ch = __ldg(reinterpret_cast<const int2 *>(np + 3));
                        h0 = box_child(A.x, A.y, A.z, A.w, B.x, B.y, R, tcull, t0);
                        h1 = box_child(B.z, B.w, C.x, C.y, C.z, C.w, R, tcull, t1);
                    } else {
                        const float4 *np = reinterpret_cast<const float4 *>(S.nodes + cur);
                        float4 s0 = __ldg(np), s1 = __ldg(np + 1), p0 = __ldg(np + 2), p1 = __ldg(np + 3);
                        float4 m4 = __ldg(np + 4);
                        ch = make_int2(__float_as_int(m4.z), __float_as_int(m4.w));
                        // a hit at exactly best.t with a smaller rank must still be found: prune only on strict >
                        h0 = cull_child(s0, p0, m4.x, c.o, c.d, inv_dd, slack, best.t, t0);
                        h1 = cull_child(s1, p1, m4.y, c.o, c.d, inv_dd, slack, best.t, t1);
                    }
                    if (COUNT) n_sph += 2;
                    const bool first0 = h0 && (t0 <= t1);            // child 0 is entered first; else child 1 if it is hit at all
                    const bool swap = h1 && !first0;
                    const int near = swap ? ch.y : ch.x;
                    const int far = swap ? ch.x : ch.y;
                    const bool both = h0 && h1, any = h0 || h1;
#if RT_STACK_TOP_REG
                    if (both) { *sptr++ = top; top = far; }          // depth <= RT_STACK_MAX - 2 is guaranteed by the build
                    cur = any ? near : top;
                    if (!any) top = *--sptr;
#else
                    if (both) *sptr++ = far;
                    if (any) cur = near; else cur = *--sptr;
#endif
                }
                nm = __ballot_sync(FULL, cur >= 0);
                if (__popc(nm) < keep) break;
            }
        }
        if (live) {
            // ---- cluster (leaf): linear scan like IntersectRayMesh (raytracer.cpp:136-154), <= RT_LEAF_MAX triangles ----
            if (cur < 0 && cur != RT_DONE) {
                uint32_t first = leaf_first(cur), cnt = leaf_count(cur);
                if (COUNT) n_clu += 1;
                // The first record word (normal + a.x) of every triangle of the cluster is requested before the first test: one
                // memory latency per cluster instead of one per triangle, and those loads pull in most 32-byte sectors of the
                // cluster's 48-byte records, so the remaining words of a front-facing triangle are mostly L1 hits.
                // Arithmetic = tri_test, same order.
                const float4 *tp = reinterpret_cast<const float4 *>(S.tris + first);
                float4 r0[RT_LEAF_MAX];
#pragma unroll
                for (int k = 0; k < RT_LEAF_MAX; ++k) if (k < (int)cnt) r0[k] = __ldg(tp + 3 * k);
#pragma unroll
                for (int k = 0; k < RT_LEAF_MAX; ++k) {
                    if (k < (int)cnt) {
                        f3 n = mk3(r0[k].x, r0[k].y, r0[k].z);
                        float dd = dot3(c.qp, n);
                        if (!(dd <= 0.0f)) {                                        // raytracer.cpp:97-98
                            float4 r1 = __ldg(tp + 3 * k + 1), r2 = __ldg(tp + 3 * k + 2);
                            f3 ap = c.o - mk3(r0[k].w, r1.x, r1.y);
                            float t = dot3(ap, n);
                            if (!(t < 0.0f)) {                                      // raytracer.cpp:102-103
                                f3 e = cross3(c.qp, ap);
                                float v = dot3(mk3(r2.y, r2.z, r2.w), e);
                                if (!(v < 0.0f || v > dd)) {                        // raytracer.cpp:108-109
                                    float w = -dot3(mk3(r1.z, r1.w, r2.x), e);
                                    if (!(w < 0.0f || (v + w) > dd)) {              // raytracer.cpp:110-111
                                        float ood = 1.0f / dd;
                                        t = t * ood; v = v * ood; w = w * ood;
                                        if (t <= best.t && t < FLT_MAX) {           // t < FLT_MAX: raytracer.cpp:149/220 against { FLT_MAX }
                                            uint32_t ti = first + (uint32_t)k;
                                            uint32_t rk = __ldg(S.tri_rank + ti);
                                            if (t < best.t || rk < best_rank) {
                                                best.t = t; best.v = v; best.w = w; best.tri = (int32_t)ti; best_rank = rk;
                                                tcull = t * 1.00001f;
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                if ((kind & 3) == 1 && best.tri >= 0) cur = RT_DONE;                     // occlusion only needs TraceRay's bool (raytracer.cpp:385)
#if RT_STACK_TOP_REG
                else { cur = top; top = *--sptr; }
#else
                else cur = *--sptr;
#endif
            }
            if (cur == RT_DONE) {
                if (kind == 0) {
                    W.hits[out_idx] = best;
                } else {
                    const float4 r = W.rad[out_idx];                                     // w: squared light distance (point light)
                    const uint32_t light = (uint32_t)kind >> 2;
                    bool lit = best.tri < 0 || ((kind & 3) == 2 && best.t * best.t <= r.w);   // raytracer.cpp:385 / 395-396
                    if (lit) {
                        uint32_t slot = __float_as_uint(W.shadow_o[out_idx].w);
                        float4 *dst = light == 0 ? W.acc + slot : W.acc_extra + (size_t)(light - 1) * W.shadow_stride + slot;
                        float4 a = *dst;
                        a.x += r.x; a.y += r.y; a.z += r.z;
                        *dst = a;
                    }
                }
                live = false;
            }
        }
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) { n_sph += __shfl_down_sync(FULL, n_sph, o); n_clu += __shfl_down_sync(FULL, n_clu, o); }
        if (lane == 0) { atomicAdd(&counters->sphere_checks, n_sph); atomicAdd(&counters->cluster_checks, n_clu); }
    }
}

// brute force over every triangle: test-only cross-check of the hierarchy's conservativeness
__global__ void k_trace_brute(DevScene S, float bias, RayQueue q, uint32_t n, HitRec *hits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 o4 = q.o[i], d4 = q.d[i];
    RayCtx c;
    c.d = mk3(d4);
    c.o = mk3(o4) + c.d * bias;
    f3 qq = c.o + c.d;
    c.qp = c.o - qq;
    HitRec best; best.t = FLT_MAX; best.v = 0; best.w = 0; best.tri = -1;
    uint32_t best_rank = 0xFFFFFFFFu;
    for (uint32_t ti = 0; ti < S.n_tris; ++ti) {
        TriRec r = S.tris[ti];
        float t, v, w;
        if (tri_test(r, c, t, v, w) && t <= best.t && t < FLT_MAX) {
            uint32_t rk = S.tri_rank[ti];
            if (t < best.t || rk < best_rank) { best.t = t; best.v = v; best.w = w; best.tri = (int32_t)ti; best_rank = rk; }
        }
    }
    hits[i] = best;
}
