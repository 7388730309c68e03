// rt_render.cu -- the per-render pool, the wave scheduler and the entry points that replace Render / RenderTask / RenderPixel /
// TraceRayColor / TraceRay. Every hit, ray and colour is computed by the kernels in rt_trace.cuh / rt_shade.cuh.
#include "rt_internal.h"
#include "rt_rng.cuh"
#include "rt_raygen.cuh"
#include "rt_shade.cuh"
#include "rt_trace.cuh"

#define RT_PI32 (3.1415927f)                                   // brt.h:23

static float host_radical_inverse(uint32_t bits) {             // raytracer.cpp:273-282
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)(bits * 2.3283064365386963e-10);
}

static void host_phong_dirs(const std::vector<float> &spec_intensity, uint32_t ss, std::vector<float4> &out) {   // raytracer.cpp:290-300
    out.resize(spec_intensity.size() * (size_t)std::max(1u, ss));
    for (size_t m = 0; m < spec_intensity.size(); ++m) {
        for (uint32_t s = 0; s < ss; ++s) {
            volatile float xi_x = (float)s / (float)ss;
            volatile float xi_y = host_radical_inverse(s);
            volatile float phi = 2.0f * RT_PI32 * xi_x;
            volatile float cp = cosf(phi);
            volatile float sp = sinf(phi);
            volatile float ct = powf(1.0f - xi_y, 1.0f / (spec_intensity[m] + 1.0f));
            volatile float st = sqrtf(1.0f - (ct * ct));
            out[m * ss + s] = make_float4(cp * st, sp * st, ct, 0.0f);
        }
    }
}


// persistent-grid sizes: resident blocks of the whole chip for the trace kernel of the selected child bound and for k_logic
int rt_render_configure(rt_scene *sc) {
    int per_sm = 0;
    if (sc->bounds == RT_BOUNDS_QBOX4) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wave<false, RT_BOUNDS_QBOX4>, RT_TRACE_BLOCK, 0));
    else if (sc->bounds == RT_BOUNDS_QBOX) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wave<false, RT_BOUNDS_QBOX>, RT_TRACE_BLOCK, 0));
    else if (sc->bounds == RT_BOUNDS_BOX) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wave<false, RT_BOUNDS_BOX>, RT_TRACE_BLOCK, 0));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wave<false, RT_BOUNDS_SPHERE>, RT_TRACE_BLOCK, 0));
    sc->trace_grid = sc->sm_count * std::max(1, per_sm);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_logic, 128, 0));
    sc->logic_grid = sc->sm_count * std::max(1, per_sm) * 2;
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// pool
// ---------------------------------------------------------------------------------------------
// Path slots in flight at once. A frame is rendered in batches of this many (pixel, sample) slots, and every batch pays its ~8 small tail
// waves; B200 has the HBM to keep whole frames in flight: 2^27 slots (~49 GB at bounce depth 2, allocated on demand, only as many as the
// render needs) run a 1080p x 128 spp frame in 2 batches instead of 8 -- measured 2^25 / 2^26 / 2^27 / 2^28: 167.1 / 163.4 / 162.8 / 162.1 ms
// (config 3), 1,581 / 1,562 / 1,557 / 1,571 ms (config 4). Never more than 40 % of the device memory that is free at the first render.
static uint32_t pool_limit(const rt_scene *sc, uint32_t depth) {
    const char *e = getenv("RT_B200_POOL");
    uint32_t v = e ? (uint32_t)strtoul(e, nullptr, 10) : 0;
    if (v) return v;
    if (sc->pool_limit_cached) return std::max(sc->pool_limit_cached, sc->pool.capacity);      // cudaMemGetInfo is a driver round trip (sporadically 10+ ms): once per scene
    uint64_t limit = 1ull << 27;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const uint64_t lights = std::max(1u, sc->n_lights);
        const uint64_t per_slot = 16 + 8 + 16 + 16 + 16ull * RT_FRAME_F4 * std::max(1u, depth) + 4 + 64 + 16 + 32 * lights + 16 * (lights - 1);
        // memory already held by this scene's pool is reusable: count it as free
        const uint64_t held = (uint64_t)sc->pool.capacity * per_slot;
        const uint64_t fit = (uint64_t)((free_b + held) * 0.4) / per_slot;
        limit = std::min(limit, std::max<uint64_t>(fit, 1ull << 20));
    } else (void)cudaGetLastError();
    const_cast<rt_scene *>(sc)->pool_limit_cached = (uint32_t)limit;
    return (uint32_t)std::max<uint64_t>(limit, sc->pool.capacity);
}

static int ensure_pool(rt_scene *sc, uint32_t capacity, uint32_t depth) {
    Pool &p = sc->pool;
    uint32_t lights = std::max(1u, sc->n_lights);
    depth = std::max(1u, depth);
    if (p.capacity >= capacity && p.depth >= depth && p.lights >= lights) return RT_OK;
    CK(cudaStreamSynchronize(sc->stream));
    p.mem.release();
    capacity = std::max(capacity, p.capacity); depth = std::max(depth, p.depth);
    p.capacity = p.depth = 0;
    size_t c = capacity;
#if RT_POOL_AOS
    CK(p.mem.alloc(&p.paths.node_T, 2 * c)); p.paths.rng_cx = nullptr;          // one 32-byte record per slot: throughput + generator state (rt_types.cuh)
#else
    CK(p.mem.alloc(&p.paths.node_T, c)); CK(p.mem.alloc(&p.paths.rng_cx, c));
#endif
    CK(p.mem.alloc(&p.paths.rng_seed, c)); CK(p.mem.alloc(&p.paths.acc, c));
    CK(p.mem.alloc(&p.paths.frames, c * RT_FRAME_F4 * depth));
    p.paths.depth = depth;
    CK(p.mem.alloc(&p.ray_cnt, c));
    p.paths.ray_cnt = nullptr;           // switched on only by the adaptive loop
    p.paths.capacity = capacity;
    for (int k = 0; k < 2; ++k) { CK(p.mem.alloc(&p.q[k].o, c)); CK(p.mem.alloc(&p.q[k].d, c)); }
    CK(p.mem.alloc(&p.hits, c));
    CK(p.mem.alloc(&p.shadow.o, c * lights)); CK(p.mem.alloc(&p.shadow.rad, c * lights));
    CK(p.mem.alloc(&p.counts, 2 + lights)); CK(p.mem.alloc(&p.totals, 1)); CK(p.mem.alloc(&p.tcount, 1));
    CK(p.mem.alloc(&p.acc_extra, c * (lights - 1))); CK(p.mem.alloc(&p.next, 1));
    CK(cudaMemsetAsync(p.acc_extra, 0, std::max<size_t>(1, c * (lights - 1)) * sizeof(float4), sc->stream));
    p.shadow.count = p.counts + 2;
    p.shadow.capacity = capacity;
    if (p.h_counts) { cudaFreeHost(p.h_counts); p.h_counts = nullptr; }
    CK(cudaMallocHost((void **)&p.h_counts, 4 * (2 + lights) * sizeof(uint32_t)));
    for (int k = 0; k < 4; ++k) if (!p.count_ev[k]) CK(cudaEventCreateWithFlags(&p.count_ev[k], cudaEventDisableTiming));
    p.capacity = capacity; p.depth = depth; p.lights = lights;
    return RT_OK;
}

static int ensure_spec_table(rt_scene *sc, uint32_t ss) {
    if (sc->spec_dir && sc->spec_dir_ss == ss) return RT_OK;
    std::vector<float4> tab;
    host_phong_dirs(sc->spec_intensity, ss, tab);
    CK(cudaStreamSynchronize(sc->stream));
    float4 *d;
    CK(sc->mem.alloc(&d, tab.size()));
    CK(cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(float4), cudaMemcpyHostToDevice, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    sc->spec_dir = d; sc->spec_dir_ss = ss; sc->d.spec_dir = d;
    return RT_OK;
}

template <typename T> static int grow(rt_scene *sc, T **buf, size_t *cap, size_t need) {
    if (*cap >= need && *buf) return RT_OK;
    CK(cudaStreamSynchronize(sc->stream));
    if (*buf) CK(cudaFree(*buf));
    *buf = nullptr; *cap = 0;
    size_t n = std::max<size_t>(need, 1);
    CK(cudaMalloc((void **)buf, n * sizeof(T)));
    *cap = n;
    return RT_OK;
}

static DevParams to_dev_params(const rt_params *p) {
    DevParams d;
    d.ray_bias = p->ray_bias; d.reflection_samples = p->reflection_samples; d.spec_samples = p->spec_samples;
    d.bounce_depth = p->bounce_depth; d.bg[0] = p->background_color[0]; d.bg[1] = p->background_color[1];
    d.bg[2] = p->background_color[2]; d.pad = 0; d.base_seed = p->base_seed;
    return d;
}

static int check_params(const rt_params *p) {
    if (!p) return fail(RT_ERR_ARG, "params is null");
    if (p->bounce_depth > 200) return fail(RT_ERR_ARG, "bounce_depth %u > 200", p->bounce_depth);
    if ((uint64_t)p->reflection_samples + p->spec_samples > 1000000u) return fail(RT_ERR_ARG, "too many reflection samples");
    // The per-path draw counter is 16 bits wide (pack_state, rt_shade.cuh). Worst case of one sample: 2 jitter draws, then every node of
    // the recursion tree (reflection + specular children, + 1 translucent continuation, bounce_depth levels below the root) takes one
    // Russian-roulette draw and one draw per diffuse child (raytracer.cpp:416-420, 519-520). Reject what could wrap the counter.
    {
        const uint64_t fan = (uint64_t)p->reflection_samples + p->spec_samples + 1u, per_node = 1u + (uint64_t)p->reflection_samples;
        uint64_t nodes = 1, level = 1, draws = 2;
        bool ok = true;
        for (uint32_t d = 0; d < p->bounce_depth && ok; ++d) {
            if (level > 65535u / fan) { ok = false; break; }
            level *= fan; nodes += level;
            if (nodes > 65535u) ok = false;
        }
        if (ok) { draws += nodes * per_node; if (draws > 65535u) ok = false; }
        if (!ok) return fail(RT_ERR_ARG, "bounce_depth %u with %u + %u samples per bounce can take more than 65535 random draws per sample (the per-path draw counter is 16 bits)",
                             p->bounce_depth, p->reflection_samples, p->spec_samples);
    }
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// the wave loop: runs every live path of the pool to completion
// ---------------------------------------------------------------------------------------------
struct WaveCfg { bool count; };

static int wave_event(rt_scene *sc, bool on) {
    if (!on) return RT_OK;
    if (sc->tev_used == sc->tev.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        sc->tev.push_back(e);
    }
    CK(cudaEventRecord(sc->tev[sc->tev_used++], sc->stream));
    return RT_OK;
}

// sums the per-wave event intervals recorded since tev_used was reset (stream must be idle)
static int collect_wave_times(rt_scene *sc) {
    const bool log = getenv("RT_B200_WAVE_LOG") != nullptr;
    for (size_t i = 0; i + 3 < sc->tev_used; i += 4) {
        float a = 0, b = 0, c = 0;
        CK(cudaEventElapsedTime(&a, sc->tev[i], sc->tev[i + 1]));
        CK(cudaEventElapsedTime(&b, sc->tev[i + 1], sc->tev[i + 2]));
        CK(cudaEventElapsedTime(&c, sc->tev[i + 2], sc->tev[i + 3]));
        sc->stats.trace_ms += a; sc->stats.logic_ms += b; sc->stats.shadow_ms += c;
        if (log && i / 4 < sc->wave_log.size()) {
            auto &wl = sc->wave_log[i / 4];
            fprintf(stderr, "[wave %3zu] closest %9llu shadow %9llu  trace %7.3f ms (%6.0f Mrays/s)  logic %7.3f ms\n", i / 4, (unsigned long long)wl.first,
                    (unsigned long long)wl.second, a, a > 0 ? (wl.first + wl.second) / a / 1e3 : 0.0, b);
        }
    }
    sc->tev_used = 0;
    sc->wave_log.clear();
    return RT_OK;
}

static uint32_t env_knob(const char *name, uint32_t dflt) {
    const char *e = getenv(name);
    uint32_t v = e ? (uint32_t)atoi(e) : dflt;
    return (v < 1 || v > 32) ? dflt : v;
}
static void set_fetch_knobs(WaveQueues &w) {
    uint32_t a = 0, b = 0, c = 0, d = 0;      // read per launch (microseconds): lets one process sweep the knobs
    { a = env_knob("RT_B200_FETCH_MIN", RT_FETCH_MIN); b = env_knob("RT_B200_FETCH_PRIMARY", RT_FETCH_PRIMARY); c = env_knob("RT_B200_FETCH_SHADOW", RT_FETCH_MIN);
              d = env_knob("RT_B200_LEAF_WAIT", RT_LEAF_WAIT); }
    w.fetch_min = a; w.fetch_min_primary = b; w.fetch_min_shadow = c; w.leaf_wait = d;
}

static WaveQueues wave_queues(rt_scene *sc, int cur, uint32_t n_closest_max) {
    Pool &p = sc->pool;
    WaveQueues w;
    w.closest = p.q[cur]; w.n_closest = p.counts + cur; w.closest_max = n_closest_max; w.hits = p.hits;
    w.shadow_o = p.shadow.o; w.shadow_dir = nullptr; w.rad = p.shadow.rad; w.n_shadow = p.shadow.count; w.shadow_stride = p.shadow.capacity;
    w.n_lights = sc->n_lights; w.acc = p.paths.acc; w.acc_extra = p.acc_extra; w.next = p.next;
    set_fetch_knobs(w);
    return w;
}

static PrimaryGen no_gen() { PrimaryGen g; memset(&g, 0, sizeof(g)); return g; }

static int launch_trace_wave(rt_scene *sc, float bias, const WaveQueues &w, const PrimaryGen &gen, uint64_t work_bound, bool count, TraceCounters *tc) {
    cudaStream_t st = sc->stream;
    CK(cudaMemsetAsync(w.next, 0, 4, st));
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sc->trace_grid, std::max<uint64_t>(1, (work_bound + RT_TRACE_BLOCK - 1) / RT_TRACE_BLOCK));
#define RT_TRACE_LAUNCH(B) do { if (count) k_trace_wave<true, B><<<grid, RT_TRACE_BLOCK, 0, st>>>(sc->d, bias, w, gen, tc); \
                                else k_trace_wave<false, B><<<grid, RT_TRACE_BLOCK, 0, st>>>(sc->d, bias, w, gen, tc); } while (0)
    if (sc->bounds == RT_BOUNDS_QBOX4) RT_TRACE_LAUNCH(RT_BOUNDS_QBOX4);
    else if (sc->bounds == RT_BOUNDS_QBOX) RT_TRACE_LAUNCH(RT_BOUNDS_QBOX);
    else if (sc->bounds == RT_BOUNDS_BOX) RT_TRACE_LAUNCH(RT_BOUNDS_BOX);
    else RT_TRACE_LAUNCH(RT_BOUNDS_SPHERE);
#undef RT_TRACE_LAUNCH
    CKL("k_trace_wave");
    return RT_OK;
}

// Runs every live path of the pool to completion. Wave w: ONE trace launch (the pending nodes' closest-hit rays +
// the shadow rays the previous shading step queued), then the shading / bounce-generation step.
//
// The host never stalls the GPU: queue sizes live in device memory (kernels read them there), and wave w + 1 is
// enqueued -- with launch bounds taken from the newest counts the host already has -- BEFORE the host waits for
// wave w's 16-byte count read-back. The read-backs only decide when to stop and feed the statistics.
#define RT_COUNT_RING 4
static int run_waves(rt_scene *sc, const DevParams &prm, uint32_t n_first, uint32_t flags, uint64_t *launches, const PrimaryGen *first_gen = nullptr) {
    const bool timed = (flags & RT_FLAG_TIME_KERNELS) != 0;
    Pool &p = sc->pool;
    cudaStream_t st = sc->stream;
    const bool count = (flags & RT_FLAG_COUNTERS) != 0;
    const uint32_t L = sc->n_lights;
    const uint32_t stride = 2 + L;                      // words per read-back slot
    if (L) CK(cudaMemsetAsync(p.counts + 2, 0, 4 * L, st));

    uint32_t bound = n_first;                           // upper bound of the closest-hit queue of the wave being issued
    auto issue = [&](uint32_t w) -> int {
        const int cur = (int)(w & 1u);
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        const PrimaryGen gen = (w == 0 && first_gen) ? *first_gen : no_gen();     // wave 0 of a render: rays are generated in place
        { int rc_ = launch_trace_wave(sc, prm.ray_bias, wave_queues(sc, cur, bound), gen, (uint64_t)bound * (1 + L), count, p.tcount); if (rc_) return rc_; }
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        CK(cudaMemsetAsync(p.counts + (cur ^ 1), 0, 4, st));
        if (L) CK(cudaMemsetAsync(p.counts + 2, 0, 4 * L, st));
        k_logic<<<std::min(cdiv(std::max(1u, bound), RT_LOGIC_CHUNK), (uint32_t)sc->logic_grid), 128, 0, st>>>(sc->d, prm, p.paths, p.q[cur], p.hits, p.counts + cur, bound, p.q[cur ^ 1], p.counts + (cur ^ 1), p.shadow, gen);
        CKL("k_logic");
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        { int rc_ = wave_event(sc, timed); if (rc_) return rc_; }
        CK(cudaMemcpyAsync(p.h_counts + (size_t)(w % RT_COUNT_RING) * stride, p.counts, 4 * stride, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(p.count_ev[w % RT_COUNT_RING], st));
        *launches += 2;
        return RT_OK;
    };

    uint64_t sh_in = 0;                                 // shadow rays traced by the wave whose read-back we wait for
    uint32_t c_in = n_first;                            // closest rays traced by that wave
    { int rc_ = issue(0); if (rc_) return rc_; }
    for (uint32_t w = 0;; ++w) {
        { int rc_ = issue(w + 1); if (rc_) return rc_; }                       // run ahead by one wave
        CK(cudaEventSynchronize(p.count_ev[w % RT_COUNT_RING]));
        const uint32_t *hc = p.h_counts + (size_t)(w % RT_COUNT_RING) * stride;
        const int cur = (int)(w & 1u);
        uint32_t c_out = hc[cur ^ 1];
        uint64_t sh_out = 0;
        for (uint32_t l = 0; l < L; ++l) sh_out += hc[2 + l];
        sc->stats.closest_rays += c_in;
        sc->stats.shadow_rays += sh_in;
        sc->stats.waves += 1;
        if (timed) sc->wave_log.push_back(std::make_pair((uint64_t)c_in, sh_in));
        c_in = c_out; sh_in = sh_out;
        bound = c_out;                                  // queue sizes never grow: every live path emits at most one ray per wave
        if (c_out == 0 && sh_out == 0) break;           // the wave already in flight finds empty queues and does nothing
    }
    if (timed) sc->wave_log.push_back(std::make_pair((uint64_t)0, (uint64_t)0));      // the run-ahead wave that found empty queues
    if (L > 1) {
        k_fold_light_acc<<<cdiv(n_first, 256), 256, 0, st>>>(p.paths.acc, p.acc_extra, n_first, p.shadow.capacity, L - 1);
        CKL("k_fold_light_acc");
        *launches += 1;
    }
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// rt_render_device / rt_render
// ---------------------------------------------------------------------------------------------
static DevCamera to_dev_camera(const rt_camera *c) {
    DevCamera d;
    static_assert(sizeof(DevCamera) == sizeof(rt_camera), "camera layout");
    memcpy(&d, c, sizeof(d));
    return d;
}

// ids_on_device: pixel_ids is a DEVICE pointer to a list that was validated when it was built (rt_comm's cached tile partition): no
// per-frame check and no per-frame upload of the list.
static int render_impl(rt_scene *sc, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                       const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                       uint32_t sample_count, uint32_t flags, float *out_dev, cudaStream_t user_stream, rt_counters *out_counters,
                       bool ids_on_device = false) {
    if (!sc || !cam || !out_dev) return fail(RT_ERR_ARG, "null argument");
    int rc = check_params(params);
    if (rc) return rc;
    if (!width || !height) return fail(RT_ERR_ARG, "empty frame");
    const uint64_t frame = (uint64_t)width * height;
    if (!pixel_ids && (uint64_t)pixel_begin + pixel_count > frame) return fail(RT_ERR_ARG, "pixel range exceeds the frame");
    if (pixel_ids && !ids_on_device) for (uint32_t k = 0; k < pixel_count; ++k) if (pixel_ids[k] >= frame) return fail(RT_ERR_ARG, "pixel id %u out of range", pixel_ids[k]);
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    // Order after the caller's stream. Handle 0 is the LEGACY default stream (what torch's default stream is): the scene stream is
    // non-blocking, so without this wait the render would not be ordered after the caller's memsets / previous reduce on it.
    if (!user_stream) user_stream = cudaStreamLegacy;
    if (user_stream != st) {
        CK(cudaEventRecord(sc->ev1, user_stream));
        CK(cudaStreamWaitEvent(st, sc->ev1, 0));
    }
    if ((flags & RT_FLAG_ADAPTIVE) && params->min_samples < params->max_samples) {
        if (params->min_samples == 0) return fail(RT_ERR_ARG, "adaptive sampling needs min_samples >= 1");
        if (flags & RT_OUT_SUM) return fail(RT_ERR_ARG, "adaptive sampling resolves per pixel; RT_OUT_SUM is not meaningful");
        sample_count = params->min_samples;        // first loop of RenderPixel (main.cpp:237-243); the second follows below
    }
    memset(&sc->stats, 0, sizeof(sc->stats));
    sc->tev_used = 0;
    uint64_t launches = 0;
    if (pixel_count == 0 || sample_count == 0) { if (out_counters) memset(out_counters, 0, sizeof(*out_counters)); return RT_OK; }

    DevParams prm = to_dev_params(params);
    DevCamera dcam = to_dev_camera(cam);
    rc = ensure_spec_table(sc, params->spec_samples);
    if (rc) return rc;
    const uint32_t limit = pool_limit(sc, params->bounce_depth);
    const uint32_t spp_chunk = std::min(sample_count, limit);
    const uint32_t pix_per_batch = std::max(1u, std::min(pixel_count, limit / spp_chunk));
    uint64_t pool_want = (uint64_t)pix_per_batch * spp_chunk;
    if ((flags & RT_FLAG_ADAPTIVE) && params->min_samples < params->max_samples)      // room for the second loop's chunks of up to 32 samples per pixel
        pool_want = std::max<uint64_t>(pool_want, (uint64_t)pixel_count * std::min(32u, params->max_samples - params->min_samples));
    rc = ensure_pool(sc, (uint32_t)std::min<uint64_t>(pool_want, (uint64_t)limit), params->bounce_depth);
    if (rc) return rc;
    Pool &p = sc->pool;

    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    uint32_t *d_ids = nullptr;
    rc = grow(sc, &sc->accum, &sc->accum_cap, pixel_count);
    if (rc) return done(rc);
    float4 *accum = sc->accum;
    CKR(cudaMemsetAsync(accum, 0, (size_t)pixel_count * sizeof(float4), st));
    if (pixel_ids && ids_on_device) d_ids = const_cast<uint32_t *>(pixel_ids);
    else if (pixel_ids) {
        rc = grow(sc, &sc->ids, &sc->ids_cap, pixel_count);
        if (rc) return done(rc);
        d_ids = sc->ids;
        CKR(cudaMemcpyAsync(d_ids, pixel_ids, (size_t)pixel_count * 4, cudaMemcpyHostToDevice, st));
        sc->stats.h2d_bytes += (uint64_t)pixel_count * 4;
    }
    CKR(cudaMemsetAsync(p.totals, 0, sizeof(WaveTotals), st));
    CKR(cudaMemsetAsync(p.tcount, 0, sizeof(TraceCounters), st));
    CKR(cudaEventRecord(sc->ev0, st));

    const bool adaptive = (flags & RT_FLAG_ADAPTIVE) && params->min_samples < params->max_samples;
    const uint32_t max_s = params->max_samples;
    bool adaptive_counts = false;                     // ray_count = first pass + the rays of the ACCEPTED samples of the second loop
    unsigned long long adaptive_first_pass_rays = 0, adaptive_accepted_rays = 0;
    uint32_t *d_nsamples = nullptr;
    if (adaptive) {
        rc = grow(sc, &sc->scratch, &sc->scratch_cap, (size_t)pixel_count * max_s);
        if (rc) return done(rc);
        rc = grow(sc, &sc->ad_u32, &sc->ad_u32_cap, 5 * (size_t)pixel_count + 8);
        if (rc) return done(rc);
        d_nsamples = sc->ad_u32;
    }
    for (uint32_t p0 = 0; p0 < pixel_count; p0 += pix_per_batch) {
        const uint32_t npix = std::min(pix_per_batch, pixel_count - p0);
        for (uint32_t s0 = 0; s0 < sample_count; s0 += spp_chunk) {
            const uint32_t ns = std::min(spp_chunk, sample_count - s0);
            const uint32_t n_slots = npix * ns;
            PrimaryGen gen;
            memset(&gen, 0, sizeof(gen));
            gen.cam = dcam; gen.base_seed = prm.base_seed; gen.pixel_ids = d_ids; gen.n_slots = n_slots; gen.spp = ns; gen.width = width;
            gen.pixel_begin = pixel_begin; gen.pixel_local0 = p0; gen.sample_begin = sample_begin + s0; gen.jitter_scale = 0.5f; gen.enabled = 1;
            rc = run_waves(sc, prm, n_slots, flags, &launches, &gen);
            if (rc) return done(rc);
            if (adaptive) k_resolve_scratch<<<cdiv(npix, 128), 128, 0, st>>>(p.paths.acc, npix, ns, accum, sc->scratch, pixel_count, p0, s0);
            else k_resolve<<<cdiv(npix, 128), 128, 0, st>>>(p.paths.acc, npix, ns, accum, p0);
            { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "launch of k_resolve failed: %s", cudaGetErrorString(e_))); }
            launches++;
        }
    }
    if (adaptive) {
        // RenderPixel's second loop (main.cpp:246-258), jitter x 1.0. The reference takes one more sample per pixel and iteration; here every
        // still-active pixel gets its next K samples at once (K = 2, 4, 8, ...: at most twice what the pixel ends up using) and
        // k_adaptive_update replays the reference's per-sample decisions over them. Samples past a pixel's stopping point are discarded,
        // rays included (per-path ray counts), so colours, sample counts and ray_count are those of the one-at-a-time loop.
        uint32_t *lists[2][2] = {{sc->ad_u32 + pixel_count, sc->ad_u32 + 2 * (size_t)pixel_count},
                                 {sc->ad_u32 + 3 * (size_t)pixel_count, sc->ad_u32 + 4 * (size_t)pixel_count}};
        uint32_t *d_count = sc->ad_u32 + 5 * (size_t)pixel_count;
        unsigned long long *d_rays = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(d_count + 1) + 7u) & ~(uintptr_t)7u);
        CKR(cudaMemsetAsync(d_rays, 0, 8, st));
        k_adaptive_init<<<cdiv(pixel_count, 256), 256, 0, st>>>(pixel_count, d_ids, pixel_begin, lists[0][0], lists[0][1], d_nsamples, max_s);
        launches++;
        adaptive_first_pass_rays = sc->stats.closest_rays + sc->stats.shadow_rays;
        p.paths.ray_cnt = p.ray_cnt;
        uint32_t n_active = pixel_count;
        int cur = 0;
        uint32_t K = 2;
        for (uint32_t samp = params->min_samples; samp < max_s && n_active > 0;) {
            // chunk = the doubling schedule, or whatever it takes to put ~4 M samples in flight (a 720x480 frame cannot fill the chip with less)
            const uint32_t K_fill = (uint32_t)std::min<uint64_t>(((4ull << 20) + n_active - 1) / n_active, 1u << 20);
            const uint32_t Kc = std::max(1u, std::min(std::min(std::max(K, K_fill), max_s - samp), p.capacity / std::max(1u, std::min(n_active, p.capacity))));
            const uint32_t na_max = std::max(1u, p.capacity / Kc);
            CKR(cudaMemsetAsync(d_count, 0, 4, st));
            for (uint32_t a0 = 0; a0 < n_active; a0 += na_max) {   // more active samples than pool slots: chunks
                const uint32_t na = std::min(na_max, n_active - a0);
                PrimaryGen gen;
                memset(&gen, 0, sizeof(gen));
                gen.cam = dcam; gen.base_seed = prm.base_seed; gen.pixel_ids = lists[cur][0] + a0; gen.n_slots = na * Kc; gen.spp = Kc; gen.width = width;
                gen.pixel_begin = 0; gen.pixel_local0 = 0; gen.sample_begin = sample_begin + samp; gen.jitter_scale = 1.0f; gen.enabled = 1;
                rc = run_waves(sc, prm, na * Kc, flags, &launches, &gen);
                if (rc) { p.paths.ray_cnt = nullptr; return done(rc); }
                k_adaptive_update<<<cdiv(na, 128), 128, 0, st>>>(p.paths.acc, p.ray_cnt, na, samp, Kc, max_s, lists[cur][0] + a0, lists[cur][1] + a0, accum, sc->scratch, (size_t)pixel_count,
                                                                d_nsamples, lists[cur ^ 1][0], lists[cur ^ 1][1], d_count, d_rays);
                { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { p.paths.ray_cnt = nullptr; return done(fail(RT_ERR_CUDA, "launch of k_adaptive_update failed: %s", cudaGetErrorString(e_))); } }
                launches++;
            }
            CKR(cudaMemcpyAsync(&n_active, d_count, 4, cudaMemcpyDeviceToHost, st));
            CKR(cudaStreamSynchronize(st));
            cur ^= 1;
            samp += Kc;
            K = std::min(K * 2u, 1u << 20);
        }
        p.paths.ray_cnt = nullptr;
        CKR(cudaMemcpyAsync(&adaptive_accepted_rays, d_rays, 8, cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
        adaptive_counts = true;
        sc->last_adaptive_pixels = pixel_count;
    }
    k_finalize<<<cdiv(pixel_count, 128), 128, 0, st>>>(accum, pixel_count, sample_count, d_nsamples, flags & 3u, (float4 *)out_dev, d_ids, pixel_begin);
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "launch of k_finalize failed: %s", cudaGetErrorString(e_))); }
    launches++;
    CKR(cudaEventRecord(sc->ev1, st));
    TraceCounters tc;
    CKR(cudaMemcpyAsync(&tc, p.tcount, sizeof(tc), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0;
    CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms;
    sc->stats.kernel_launches = launches;
    { int rc_ = collect_wave_times(sc); if (rc_) return done(rc_); }
    if (out_counters) {
        out_counters->ray_count = adaptive_counts ? adaptive_first_pass_rays + adaptive_accepted_rays : sc->stats.closest_rays + sc->stats.shadow_rays;
        out_counters->sphere_check_count = tc.sphere_checks;
        out_counters->mesh_check_count = tc.cluster_checks;
    }
    if (user_stream != st) {      // make the caller's stream see the result
        CKR(cudaEventRecord(sc->ev1, st));
        CKR(cudaStreamWaitEvent(user_stream, sc->ev1, 0));
    }
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_render_device(rt_scene *scene, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                                const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                                uint32_t sample_count, uint32_t flags, float *out_rgba_device, void *stream, rt_counters *out_counters) {
    g_err.clear();
    return render_impl(scene, cam, params, width, height, pixel_ids, pixel_begin, pixel_count, sample_begin, sample_count, flags,
                       out_rgba_device, (cudaStream_t)stream, out_counters);
}

// library-internal (rt_comm.cu): rt_render_device with a device-resident, pre-validated pixel list
int rt_render_device_ids(rt_scene *scene, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height, const uint32_t *pixel_ids_device,
                         uint32_t pixel_count, uint32_t sample_begin, uint32_t sample_count, uint32_t flags, float *out_rgba_device, rt_counters *out_counters) {
    g_err.clear();
    return render_impl(scene, cam, params, width, height, pixel_ids_device, 0, pixel_count, sample_begin, sample_count, flags, out_rgba_device, nullptr,
                       out_counters, true);
}

extern "C" int rt_render(rt_scene *scene, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                         const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                         uint32_t sample_count, uint32_t flags, float *out_rgba_host, rt_counters *out_counters) {
    g_err.clear();
    if (!scene || !out_rgba_host) return fail(RT_ERR_ARG, "null argument");
    if (flags & RT_OUT_FULLFRAME) return fail(RT_ERR_ARG, "RT_OUT_FULLFRAME is a device-output mode");
    CK(cudaSetDevice(scene->device));
    if ((flags & RT_FLAG_PIN_HOST) && pixel_count) scene->pinned_out.pin(out_rgba_host, (size_t)pixel_count * 16);
    flags &= ~(uint32_t)RT_FLAG_PIN_HOST;
    int rc = grow(scene, &scene->out_stage, &scene->out_stage_cap, (size_t)pixel_count * 4);
    if (rc) return rc;
    float *d_out = scene->out_stage;
    rc = render_impl(scene, cam, params, width, height, pixel_ids, pixel_begin, pixel_count, sample_begin, sample_count, flags,
                         d_out, nullptr, out_counters);
    if (rc == RT_OK && pixel_count) {
        cudaError_t e = cudaMemcpyAsync(out_rgba_host, d_out, (size_t)pixel_count * 16, cudaMemcpyDeviceToHost, scene->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(scene->stream);
        if (e != cudaSuccess) rc = fail(RT_ERR_CUDA, "framebuffer download failed: %s", cudaGetErrorString(e));
        scene->stats.d2h_bytes += (uint64_t)pixel_count * 16;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// rt_trace_rays / rt_trace_primary / rt_trace_color
// ---------------------------------------------------------------------------------------------
extern "C" int rt_trace_rays(rt_scene *sc, const rt_params *params, const rt_ray *rays, uint64_t n, int mode, rt_hit *out_hits,
                             rt_counters *out_counters) {
    g_err.clear();
    if (!sc || !params || (n && (!rays || !out_hits))) return fail(RT_ERR_ARG, "null argument");
    if (mode != RT_TRACE_CLOSEST && mode != RT_TRACE_ANY && mode != RT_TRACE_BRUTE) return fail(RT_ERR_ARG, "unknown trace mode %d", mode);
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    memset(&sc->stats, 0, sizeof(sc->stats));
    const uint32_t chunk = 1u << 20;
    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    float *d_rays; RayQueue q; HitRec *hits; ApiHit *api; TraceCounters *tc; float4 *rad, *acc; uint32_t *cnt;
    uint32_t cap = (uint32_t)std::min<uint64_t>(n ? n : 1, chunk);
    CKR(tmp.alloc(&d_rays, 6 * (size_t)cap)); CKR(tmp.alloc(&q.o, cap)); CKR(tmp.alloc(&q.d, cap)); CKR(tmp.alloc(&hits, cap));
    CKR(tmp.alloc(&api, cap)); CKR(tmp.alloc(&tc, 1)); CKR(tmp.alloc(&rad, cap)); CKR(tmp.alloc(&acc, cap)); CKR(tmp.alloc(&cnt, 4));
    CKR(cudaMemsetAsync(tc, 0, sizeof(TraceCounters), st));
    CKR(cudaEventRecord(sc->ev0, st));
    for (uint64_t b = 0; b < n; b += chunk) {
        uint32_t m = (uint32_t)std::min<uint64_t>(chunk, n - b);
        CKR(cudaMemcpyAsync(d_rays, rays + b, (size_t)m * sizeof(rt_ray), cudaMemcpyHostToDevice, st));
        WaveQueues w;
        memset(&w, 0, sizeof(w));
        w.next = cnt + 1; w.hits = hits; w.acc = acc; w.acc_extra = acc; w.rad = rad; w.shadow_o = q.o; w.shadow_dir = q.d; w.closest = q;
        w.n_shadow = cnt; w.shadow_stride = cap; set_fetch_knobs(w);
        int rc = RT_OK;
        if (mode == RT_TRACE_ANY) {
            k_rays_to_shadow_queue<<<cdiv(m, 256), 256, 0, st>>>(d_rays, m, q.o, q.d, rad, acc, cnt);
            w.closest_max = 0; w.n_lights = 1;
            rc = launch_trace_wave(sc, params->ray_bias, w, no_gen(), m, true, tc);
            k_occlusion_to_api<<<cdiv(m, 256), 256, 0, st>>>(acc, m, api);
        } else {
            k_upload_rays<<<cdiv(m, 256), 256, 0, st>>>(d_rays, m, q);
            w.closest_max = m; w.n_lights = 0;
            if (mode == RT_TRACE_BRUTE) k_trace_brute<<<cdiv(m, 128), 128, 0, st>>>(sc->d, params->ray_bias, q, m, hits);
            else rc = launch_trace_wave(sc, params->ray_bias, w, no_gen(), m, true, tc);
            k_hits_to_api<<<cdiv(m, 256), 256, 0, st>>>(sc->d, params->ray_bias, q, hits, m, api);
        }
        if (rc) return done(rc);
        { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "trace launch failed: %s", cudaGetErrorString(e_))); }
        CKR(cudaMemcpyAsync(out_hits + b, api, (size_t)m * sizeof(ApiHit), cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
        sc->stats.kernel_launches += 3;
        if (mode == RT_TRACE_ANY) sc->stats.shadow_rays += m; else sc->stats.closest_rays += m;
    }
    CKR(cudaEventRecord(sc->ev1, st));
    TraceCounters htc;
    CKR(cudaMemcpyAsync(&htc, tc, sizeof(htc), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0; CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms;
    if (out_counters) { out_counters->ray_count = n; out_counters->sphere_check_count = htc.sphere_checks; out_counters->mesh_check_count = htc.cluster_checks; }
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_trace_primary(rt_scene *sc, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                                const uint32_t *pixel_ids, uint32_t pixel_begin, uint32_t pixel_count, uint32_t sample_begin,
                                uint32_t sample_count, rt_ray *out_rays, rt_hit *out_hits) {
    g_err.clear();
    if (!sc || !cam) return fail(RT_ERR_ARG, "null argument");
    int rc = check_params(params);
    if (rc) return rc;
    if (!width || !height) return fail(RT_ERR_ARG, "empty frame");
    const uint64_t frame = (uint64_t)width * height;
    if (!pixel_ids && (uint64_t)pixel_begin + pixel_count > frame) return fail(RT_ERR_ARG, "pixel range exceeds the frame");
    if (pixel_ids) for (uint32_t k = 0; k < pixel_count; ++k) if (pixel_ids[k] >= frame) return fail(RT_ERR_ARG, "pixel id %u out of range", pixel_ids[k]);
    if (!pixel_count || !sample_count) return RT_OK;
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    memset(&sc->stats, 0, sizeof(sc->stats));
    const uint32_t limit = 1u << 20;
    const uint32_t spp_chunk = std::min(sample_count, limit);
    const uint32_t pix_per_batch = std::max(1u, std::min(pixel_count, limit / spp_chunk));
    rc = ensure_pool(sc, pix_per_batch * spp_chunk, params->bounce_depth);
    if (rc) return rc;
    Pool &p = sc->pool;
    DevParams prm = to_dev_params(params);
    DevCamera dcam = to_dev_camera(cam);
    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    uint32_t *d_ids = nullptr; float *d_rays; ApiHit *api;
    const uint32_t cap = pix_per_batch * spp_chunk;
    CKR(tmp.alloc(&d_rays, 6 * (size_t)cap)); CKR(tmp.alloc(&api, cap));
    if (pixel_ids) { CKR(tmp.alloc(&d_ids, pixel_count)); CKR(cudaMemcpyAsync(d_ids, pixel_ids, (size_t)pixel_count * 4, cudaMemcpyHostToDevice, st)); }
    CKR(cudaEventRecord(sc->ev0, st));
    for (uint32_t p0 = 0; p0 < pixel_count; p0 += pix_per_batch) {
        const uint32_t npix = std::min(pix_per_batch, pixel_count - p0);
        for (uint32_t s0 = 0; s0 < sample_count; s0 += spp_chunk) {
            const uint32_t ns = std::min(spp_chunk, sample_count - s0);
            const uint32_t m = npix * ns;
            PrimaryGen gen;
            memset(&gen, 0, sizeof(gen));
            gen.cam = dcam; gen.base_seed = prm.base_seed; gen.pixel_ids = d_ids; gen.n_slots = m; gen.spp = ns; gen.width = width;
            gen.pixel_begin = pixel_begin; gen.pixel_local0 = p0; gen.sample_begin = sample_begin + s0; gen.jitter_scale = 0.5f; gen.enabled = 1;
            k_raygen<<<cdiv(m, 256), 256, 0, st>>>(gen, prm, p.paths, p.q[0], p.counts);
            if (out_hits) {
                WaveQueues w = wave_queues(sc, 0, m);
                w.n_closest = nullptr; w.n_lights = 0;
                rc = launch_trace_wave(sc, prm.ray_bias, w, no_gen(), m, false, p.tcount);
                if (rc) return done(rc);
                k_hits_to_api<<<cdiv(m, 256), 256, 0, st>>>(sc->d, prm.ray_bias, p.q[0], p.hits, m, api);
                sc->stats.kernel_launches += 2; sc->stats.closest_rays += m;
            }
            if (out_rays) k_queue_to_rays<<<cdiv(m, 256), 256, 0, st>>>(p.q[0], m, d_rays);
            { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "primary launch failed: %s", cudaGetErrorString(e_))); }
            sc->stats.kernel_launches += 1;
            // entries are pixel-major over the WHOLE call: (p0 + k) * sample_count + s0 + s
            for (uint32_t k = 0; k < npix; ++k) {
                size_t dst = (size_t)(p0 + k) * sample_count + s0, src = (size_t)k * ns;
                if (out_hits) CKR(cudaMemcpyAsync(out_hits + dst, api + src, (size_t)ns * sizeof(ApiHit), cudaMemcpyDeviceToHost, st));
                if (out_rays) CKR(cudaMemcpyAsync(out_rays + dst, d_rays + 6 * src, (size_t)ns * sizeof(rt_ray), cudaMemcpyDeviceToHost, st));
                if (ns == sample_count) {   // contiguous: one copy covers the whole batch
                    if (out_hits) CKR(cudaMemcpyAsync(out_hits + dst, api + src, (size_t)ns * npix * sizeof(ApiHit), cudaMemcpyDeviceToHost, st));
                    if (out_rays) CKR(cudaMemcpyAsync(out_rays + dst, d_rays + 6 * src, (size_t)ns * npix * sizeof(rt_ray), cudaMemcpyDeviceToHost, st));
                    break;
                }
            }
            CKR(cudaStreamSynchronize(st));
        }
    }
    CKR(cudaEventRecord(sc->ev1, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0; CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms;
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_trace_color(rt_scene *sc, const rt_params *params, const rt_ray *rays, const uint64_t *seeds, uint64_t n,
                              float *out_rgba, rt_counters *out_counters) {
    g_err.clear();
    if (!sc || (n && (!rays || !seeds || !out_rgba))) return fail(RT_ERR_ARG, "null argument");
    int rc = check_params(params);
    if (rc) return rc;
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = sc->stream;
    memset(&sc->stats, 0, sizeof(sc->stats));
    if (out_counters) memset(out_counters, 0, sizeof(*out_counters));
    if (!n) return RT_OK;
    rc = ensure_spec_table(sc, params->spec_samples);
    if (rc) return rc;
    const uint32_t chunk = (uint32_t)std::min<uint64_t>(n, 1u << 20);
    rc = ensure_pool(sc, chunk, params->bounce_depth);
    if (rc) return rc;
    Pool &p = sc->pool;
    DevParams prm = to_dev_params(params);
    DevArena tmp;
    auto done = [&](int r) { tmp.release(); return r; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); } while (0)
    float *d_rays; uint64_t *d_seeds;
    CKR(tmp.alloc(&d_rays, 6 * (size_t)chunk)); CKR(tmp.alloc(&d_seeds, chunk));
    CKR(cudaMemsetAsync(p.tcount, 0, sizeof(TraceCounters), st));
    uint64_t launches = 0;
    CKR(cudaEventRecord(sc->ev0, st));
    for (uint64_t b = 0; b < n; b += chunk) {
        uint32_t m = (uint32_t)std::min<uint64_t>(chunk, n - b);
        CKR(cudaMemcpyAsync(d_rays, rays + b, (size_t)m * sizeof(rt_ray), cudaMemcpyHostToDevice, st));
        CKR(cudaMemcpyAsync(d_seeds, seeds + b, (size_t)m * 8, cudaMemcpyHostToDevice, st));
        k_paths_from_rays<<<cdiv(m, 256), 256, 0, st>>>(prm, p.paths, p.q[0], m, d_rays, d_seeds, p.counts);
        { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return done(fail(RT_ERR_CUDA, "launch of k_paths_from_rays failed: %s", cudaGetErrorString(e_))); }
        launches++;
        rc = run_waves(sc, prm, m, RT_FLAG_COUNTERS, &launches);
        if (rc) return done(rc);
        CKR(cudaMemcpyAsync(out_rgba + 4 * b, p.paths.acc, (size_t)m * 16, cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
    }
    CKR(cudaEventRecord(sc->ev1, st));
    TraceCounters tc;
    CKR(cudaMemcpyAsync(&tc, p.tcount, sizeof(tc), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    float ms = 0; CKR(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.gpu_ms = ms; sc->stats.kernel_launches = launches;
    if (out_counters) {
        out_counters->ray_count = sc->stats.closest_rays + sc->stats.shadow_rays;
        out_counters->sphere_check_count = tc.sphere_checks; out_counters->mesh_check_count = tc.cluster_checks;
    }
    return done(RT_OK);
#undef CKR
}

extern "C" int rt_rng_kat(int device, uint64_t seed, uint32_t n, uint64_t *out_host) {
    g_err.clear();
    if (!out_host) return fail(RT_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    uint64_t *d;
    CK(cudaMalloc((void **)&d, std::max(1u, n) * 8));
    k_rng_kat<<<1, 1>>>(seed, n, d);
    cudaError_t e = cudaMemcpy(out_host, d, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, "rng kat failed: %s", cudaGetErrorString(e));
    return RT_OK;
}

// Per-pixel sample counts (RenderPixel's final `samp`, main.cpp:262) of the last RT_FLAG_ADAPTIVE render on this scene.
extern "C" int rt_get_sample_counts(rt_scene *sc, uint32_t *out_host, uint32_t n) {
    g_err.clear();
    if (!sc || !out_host) return fail(RT_ERR_ARG, "null argument");
    if (n > sc->last_adaptive_pixels || !sc->ad_u32) return fail(RT_ERR_STATE, "no adaptive render of >= %u pixels on this scene", n);
    CK(cudaSetDevice(sc->device));
    CK(cudaMemcpyAsync(out_host, sc->ad_u32, (size_t)n * 4, cudaMemcpyDeviceToHost, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    return RT_OK;
}
