// rt_trace_pool.cuh -- K3p: the wave trace kernel with K rays per lane.
//
// Same job, same results as k_trace_wave (rt_trace.cuh): TraceRay (raytracer.cpp:159-232) for every closest-hit ray of the wave and
// every shadow ray of the previous shading step. What changes is how a warp keeps its lanes busy.
//
// ncu on k_trace_wave (profiles/): 85-90 % of all warp instructions are the node loop, and it runs with ~10 of 32 lanes active:
// a lane that has reached a cluster (or finished its ray) waits for the slowest descent of the warp. Here every lane owns K ray
// SLOTS instead of one ray. The lane descends with one slot; when that slot parks (reached a cluster / finished) the lane switches
// to its other slot and keeps opening nodes, so the node loop stays nearly full. Parked clusters are scanned, finished rays written
// out and idle slots refilled in bulk when enough lanes have run dry.
//
// Slot state is split by how often it is touched:
//   * per node visit (38 per ray at 10 M triangles): the 9 slab-test constants + best t -- 48 bytes per slot in SHARED memory
//     (3 x LDS.128 per slot switch; a lane works on registers while it stays on a slot);
//   * per cluster visit (~2.5 per ray): origin / direction are re-read from the ray queue the kernel was given, the hit record
//     lives in its output array from the first improvement on -- nothing of it is carried in registers;
//   * the traversal stack: per-slot in local memory, stack[s][0] = sentinel.
// Axis-aligned child bounds only (BNode). Closest-hit ties resolve by rank exactly as in k_trace_wave.
#pragma once
#include "rt_trace.cuh"

#ifndef RT_POOL_K_DEFAULT
#define RT_POOL_K_DEFAULT 1
#endif
#ifndef RT_POOL_MINB
#define RT_POOL_MINB 8
#endif
#define RT_IDLE ((int)0x80000001)      // slot holds no ray; RT_DONE: ray finished, result not yet written
RT_DEVICE bool is_cluster_ref(int cur) { return cur < 0 && cur > RT_IDLE; }

// origin / direction of the ray held by a slot, as the queues describe it (same derivation as the fetch in k_trace_wave)
RT_DEVICE void pool_ray(const DevScene &S, const WaveQueues &W, const PrimaryGen &G, uint32_t kind, uint32_t light, uint32_t out_idx, f3 &o, f3 &d) {
    if (kind == 0) {
        float4 d4 = W.closest.d[out_idx];
        d = mk3(d4);
        if (G.enabled) o = mk3(G.cam.pos[0], G.cam.pos[1], G.cam.pos[2]);            // MakeCameraRay: origin = camera position (main.cpp:175)
        else o = mk3(W.closest.o[out_idx]);
    } else {
        float4 o4 = W.shadow_o[out_idx];
        o = mk3(o4);
        if (W.shadow_dir) d = mk3(W.shadow_dir[out_idx]);
        else {
            const DevLight &Lt = S.lights[light];
            d = Lt.type == 0 ? mk3(Lt.facing[0], Lt.facing[1], Lt.facing[2]) * -1.0f                       // raytracer.cpp:240
                             : normalize3(mk3(Lt.position[0], Lt.position[1], Lt.position[2]) - o);         // raytracer.cpp:243
        }
    }
}

// LOCKSTEP = false: a lane descends with one slot and switches when it parks (fills SIMD gaps).
// LOCKSTEP = true:  every node-loop iteration advances ALL descending slots of the lane: the node loads of the K slots are issued
//                   back to back, so K dependent-load chains overlap per lane (memory-level parallelism x K at the same occupancy).
template <bool COUNT, int K, bool LOCKSTEP>
__global__ void __launch_bounds__(RT_TRACE_BLOCK, (LOCKSTEP && K == 2) ? RT_POOL_MINB : 1) k_trace_pool(DevScene S, float bias, WaveQueues W, PrimaryGen G, TraceCounters *counters) {
    extern __shared__ float4 pool_smem[];
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // slot s of this lane: st[s * 96 + 0] = (ix iy iz clx)  [+1] = (cly clz chx chy)  [+2] = (chz best_t out_idx meta); lane stride 48 B: LDS.128 conflict-free
    float4 *st = pool_smem + (size_t)warp * (K * 96) + lane * 3;
    const uint32_t nC = G.enabled ? G.n_slots : (W.n_closest ? min(*W.n_closest, W.closest_max) : W.closest_max);
    uint32_t total = nC;
    for (uint32_t l = 0; l < W.n_lights; ++l) total += min(W.n_shadow[l], W.shadow_stride);
    unsigned long long n_sph = 0, n_clu = 0;

    int curS[K], spS[K];
    int stack[K][RT_STACK_MAX];
#pragma unroll
    for (int s = 0; s < K; ++s) { curS[s] = RT_IDLE; spS[s] = 1; stack[s][0] = RT_DONE; }
    bool exhausted = false;
    uint32_t fetch_min = G.enabled ? W.fetch_min_primary : W.fetch_min;

    while (true) {
        // ---- refill: each slot set is refilled on its own, one atomicAdd per refill ----
#pragma unroll
        for (int s = 0; s < K; ++s) {
            uint32_t idle = __ballot_sync(FULL, curS[s] == RT_IDLE);
            if (!exhausted && (idle == FULL || __popc(idle) >= fetch_min)) {
                uint32_t n_idle = __popc(idle), base = 0;
                if (lane == 0) base = atomicAdd(W.next, n_idle);
                base = __shfl_sync(FULL, base, 0);
                if (base + n_idle >= total) exhausted = true;
                if (base + n_idle >= nC) fetch_min = W.fetch_min_shadow;
                if (curS[s] == RT_IDLE) {
                    uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < total) {
                        f3 o, dir; uint32_t kind, light = 0, out_idx;
                        if (idx < nC) {
                            kind = 0; out_idx = idx;
                            if (G.enabled) {
                                PathRng pr; primary_ray(G, idx, pr, o, dir);
                                W.closest.d[idx] = mk4(dir, 0.0f);               // read back at cluster visits and by the wave-0 shading step
                            } else { o = mk3(W.closest.o[idx]); dir = mk3(W.closest.d[idx]); }
                        } else {
                            uint32_t j = idx - nC;
                            while (true) { uint32_t ns = min(W.n_shadow[light], W.shadow_stride); if (j < ns) break; j -= ns; light++; }
                            size_t e = (size_t)light * W.shadow_stride + j;
                            out_idx = (uint32_t)e;
                            kind = W.rad[e].w < 0.0f ? 1u : 2u;
                            pool_ray(S, W, G, kind, light, out_idx, o, dir);
                        }
                        f3 ob = o + dir * bias;                                   // raytracer.cpp:163
                        float slack = RT_CULL_SLACK * (fabsf(ob.x) + fabsf(ob.y) + fabsf(ob.z) + S.cull_bound);
                        BoxRay R; box_ray_setup(R, ob, dir, slack);
                        st[s * 96 + 0] = make_float4(R.ix, R.iy, R.iz, R.clx);
                        st[s * 96 + 1] = make_float4(R.cly, R.clz, R.chx, R.chy);
                        st[s * 96 + 2] = make_float4(R.chz, FLT_MAX, __uint_as_float(out_idx), __uint_as_float(kind | (light << 2)));   // best.t = FLT_MAX: raytracer.cpp:166
                        curS[s] = S.n_tris ? S.root : RT_DONE; spS[s] = 1;
                    }
                }
            }
        }
        {
            bool any = false;
#pragma unroll
            for (int s = 0; s < K; ++s) any |= curS[s] != RT_IDLE;
            if (__ballot_sync(FULL, any) == 0) break;
        }

        if (LOCKSTEP) {
            // ---- node phase, lockstep: leave once `leaf_wait` slots of the warp are parked (cluster reached / ray finished) ----
            while (true) {
                bool desc = false; uint32_t parked = 0;
#pragma unroll
                for (int s = 0; s < K; ++s) { desc |= curS[s] >= 0; parked += __popc(__ballot_sync(FULL, curS[s] < 0 && curS[s] != RT_IDLE)); }
                if (__ballot_sync(FULL, desc) == 0 || parked >= W.leaf_wait) break;
                float4 nA[K], nB[K], nC[K]; int2 nch[K];
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    if (curS[s] >= 0) {
                        const float4 *np = reinterpret_cast<const float4 *>(S.bnodes + curS[s]);
                        nA[s] = __ldg(np); nB[s] = __ldg(np + 1); nC[s] = __ldg(np + 2);
                        nch[s] = __ldg(reinterpret_cast<const int2 *>(np + 3));
                    }
                }
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    if (curS[s] >= 0) {
                        float4 f0 = st[s * 96 + 0], f1 = st[s * 96 + 1], f2 = st[s * 96 + 2];
                        BoxRay R;
                        R.ix = f0.x; R.iy = f0.y; R.iz = f0.z; R.clx = f0.w; R.cly = f1.x; R.clz = f1.y; R.chx = f1.z; R.chy = f1.w; R.chz = f2.x;
                        const float tcull = f2.y * 1.00001f;
                        float t0, t1;
                        bool h0 = box_child(nA[s].x, nA[s].y, nA[s].z, nA[s].w, nB[s].x, nB[s].y, R, tcull, t0);
                        bool h1 = box_child(nB[s].z, nB[s].w, nC[s].x, nC[s].y, nC[s].z, nC[s].w, R, tcull, t1);
                        if (COUNT) n_sph += 2;
                        bool second_first = h1 & (!h0 | (t1 < t0));
                        int near = second_first ? nch[s].y : nch[s].x;
                        int far = second_first ? nch[s].x : nch[s].y;
                        int sp = spS[s];
                        if (h0 & h1) stack[s][sp++] = far;
                        curS[s] = (h0 | h1) ? near : stack[s][--sp];
                        spS[s] = sp;
                    }
                }
            }
        } else
        // ---- node phase: descend, switching slots, until `leaf_wait` lanes have no slot left to descend with ----
        {
            int sel = -1, cur = RT_IDLE, sp = 1;
            BoxRay R; float tcull = 0.0f;
            R.ix = R.iy = R.iz = R.clx = R.cly = R.clz = R.chx = R.chy = R.chz = 0.0f;
            auto pick = [&]() {
                sel = -1; cur = RT_IDLE;
#pragma unroll
                for (int s = K - 1; s >= 0; --s) if (curS[s] >= 0) { sel = s; cur = curS[s]; sp = spS[s]; }
                if (sel >= 0) {
                    float4 f0 = st[sel * 96 + 0], f1 = st[sel * 96 + 1], f2 = st[sel * 96 + 2];
                    R.ix = f0.x; R.iy = f0.y; R.iz = f0.z; R.clx = f0.w; R.cly = f1.x; R.clz = f1.y; R.chx = f1.z; R.chy = f1.w; R.chz = f2.x;
                    tcull = f2.y * 1.00001f;                                      // FLT_MAX -> inf: nothing pruned by distance
                }
            };
            auto put_back = [&]() {
#pragma unroll
                for (int s = 0; s < K; ++s) if (s == sel) { curS[s] = cur; spS[s] = sp; }
            };
            pick();
            uint32_t nm = __ballot_sync(FULL, cur >= 0);
            const int keep = max(1, __popc(nm) - (int)W.leaf_wait);
            while (nm != 0) {
                if (cur >= 0) {
                    const float4 *np = reinterpret_cast<const float4 *>(S.bnodes + cur);
                    float4 A = __ldg(np), B = __ldg(np + 1), C = __ldg(np + 2);
                    int2 ch = __ldg(reinterpret_cast<const int2 *>(np + 3));
                    float t0, t1;
                    bool h0 = box_child(A.x, A.y, A.z, A.w, B.x, B.y, R, tcull, t0);
                    bool h1 = box_child(B.z, B.w, C.x, C.y, C.z, C.w, R, tcull, t1);
                    if (COUNT) n_sph += 2;
                    bool second_first = h1 & (!h0 | (t1 < t0));
                    int near = second_first ? ch.y : ch.x;
                    int far = second_first ? ch.x : ch.y;
                    if (h0 & h1) stack[sel][sp++] = far;                          // depth <= RT_STACK_MAX - 2 is guaranteed by the build
                    cur = (h0 | h1) ? near : stack[sel][--sp];
                    if (cur < 0) { put_back(); pick(); }                          // parked: carry on with another slot of this lane
                }
                nm = __ballot_sync(FULL, cur >= 0);
                if (__popc(nm) < keep) break;
            }
            if (cur >= 0) put_back();
        }

        // ---- cluster phase: linear scan like IntersectRayMesh (raytracer.cpp:136-154), one parked slot per lane and pass ----
        for (int pass = 0; pass < K; ++pass) {
            int sel = -1;
#pragma unroll
            for (int s = K - 1; s >= 0; --s) if (is_cluster_ref(curS[s])) sel = s;
            if (__ballot_sync(FULL, sel >= 0) == 0) break;
            if (sel >= 0) {
                int cur = RT_IDLE, sp = 1;
#pragma unroll
                for (int s = 0; s < K; ++s) if (s == sel) { cur = curS[s]; sp = spS[s]; }
                float4 f2 = st[sel * 96 + 2];
                const uint32_t out_idx = __float_as_uint(f2.z), meta = __float_as_uint(f2.w), kind = meta & 3u;
                RayCtx c;
                {
                    f3 o, dir; pool_ray(S, W, G, kind, meta >> 2, out_idx, o, dir);
                    c.d = dir;
                    c.o = o + dir * bias;                                         // raytracer.cpp:163
                    f3 q = c.o + dir;
                    c.qp = c.o - q;
                }
                float bt = f2.y, bv = 0.0f, bw = 0.0f; int bti = -1;
                uint32_t brk = 0xFFFFFFFFu; bool brk_known = !(bt < FLT_MAX);     // no earlier hit: nothing to tie with
                uint32_t first = leaf_first(cur), cnt = leaf_count(cur);
                if (COUNT) n_clu += 1;
                for (uint32_t k = 0; k < cnt; ++k) {
                    uint32_t ti = first + k;
                    const float4 *tp = reinterpret_cast<const float4 *>(S.tris + ti);
                    TriRec r; r.r0 = __ldg(tp); r.r1 = __ldg(tp + 1); r.r2 = __ldg(tp + 2);
                    float t, v, w;
                    if (tri_test(r, c, t, v, w) && t <= bt && t < FLT_MAX) {        // t < FLT_MAX: raytracer.cpp:149/220 against { FLT_MAX }
                        uint32_t rk = __ldg(S.tri_rank + ti);
                        if (!(t < bt) && !brk_known) {                              // tie with a hit of an earlier cluster: fetch its rank
                            int pt = kind == 0 ? W.hits[out_idx].tri : -1;
                            brk = pt >= 0 ? __ldg(S.tri_rank + pt) : 0u;            // shadow rays: the tie cannot change the result
                            brk_known = true;
                        }
                        if (t < bt || rk < brk) { bt = t; bv = v; bw = w; bti = (int)ti; brk = rk; brk_known = true; }
                    }
                }
                if (bti >= 0) {
                    st[sel * 96 + 2].y = bt;
                    if (kind == 0) { HitRec h; h.t = bt; h.v = bv; h.w = bw; h.tri = bti; W.hits[out_idx] = h; }
                }
                cur = (kind == 1 && bt < FLT_MAX) ? RT_DONE : stack[sel][--sp];     // occlusion only needs TraceRay's bool (raytracer.cpp:385)
#pragma unroll
                for (int s = 0; s < K; ++s) if (s == sel) { curS[s] = cur; spS[s] = sp; }
            }
        }

        // ---- finished rays ----
#pragma unroll
        for (int s = 0; s < K; ++s) {
            if (curS[s] == RT_DONE) {
                float4 f2 = st[s * 96 + 2];
                const float bt = f2.y;
                const uint32_t out_idx = __float_as_uint(f2.z), meta = __float_as_uint(f2.w), kind = meta & 3u, light = meta >> 2;
                if (kind == 0) {
                    if (!(bt < FLT_MAX)) { HitRec h; h.t = FLT_MAX; h.v = 0.0f; h.w = 0.0f; h.tri = -1; W.hits[out_idx] = h; }
                } else {
                    float4 r = W.rad[out_idx];
                    bool lit = !(bt < FLT_MAX) || (kind == 2 && bt * bt <= r.w);    // raytracer.cpp:385 / 395-396
                    if (lit) {
                        uint32_t slot = __float_as_uint(W.shadow_o[out_idx].w);
                        float4 *dst = light == 0 ? W.acc + slot : W.acc_extra + (size_t)(light - 1) * W.shadow_stride + slot;
                        float4 a = *dst;
                        a.x += r.x; a.y += r.y; a.z += r.z;
                        *dst = a;
                    }
                }
                curS[s] = RT_IDLE;
            }
        }
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) { n_sph += __shfl_down_sync(FULL, n_sph, o); n_clu += __shfl_down_sync(FULL, n_clu, o); }
        if (lane == 0) { atomicAdd(&counters->sphere_checks, n_sph); atomicAdd(&counters->cluster_checks, n_clu); }
    }
}
