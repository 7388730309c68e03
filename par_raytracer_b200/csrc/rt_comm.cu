// rt_comm.cu -- Render()'s partition and the combine that replaces MPI_Gather (main.cpp:311-319, 345-347) on N GPUs.
//
// Two transports behind one interface:
//   * NCCL (one process or thread per GPU, any topology): ncclReduce(SUM) of the float4 frames to the root on the render
//     stream, then the resolve kernel the partition implies. libnccl.so.2 is bound at run time with dlopen -- the copy the
//     host process already loaded (torch's) if there is one -- so librt_b200.so itself has no NCCL link dependency.
//   * peer memory (one process, NVLink / NVSwitch): every GPU's resolve kernel stores its pixels straight into the ROOT GPU's
//     frame (tiles: no reduce and no zero-fill at all), or the root adds its peers' sample sums with one kernel that reads
//     their frames over NVLink in rank order (deterministic, unlike a ring / tree reduce whose order depends on the topology).
#include "rt_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <thread>

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("RT_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy the host process already uses (e.g. torch's)
        for (int k = 0; !h && k < 3; ++k) if (names[k]) h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
        if (!h) { const char *de = dlerror(); api.why = std::string("libnccl.so.2 not found: ") + (de ? de : ""); return; }     // dlerror() clears itself: call it once
#define RT_SYM(field, name) do { *(void **)(&api.field) = dlsym(h, name); if (!api.field) { api.why = "symbol " name " missing in libnccl"; return; } } while (0)
        RT_SYM(GetUniqueId, "ncclGetUniqueId"); RT_SYM(CommInitRank, "ncclCommInitRank"); RT_SYM(CommInitAll, "ncclCommInitAll");
        RT_SYM(CommDestroy, "ncclCommDestroy"); RT_SYM(Reduce, "ncclReduce"); RT_SYM(GetErrorString, "ncclGetErrorString");
        RT_SYM(Broadcast, "ncclBroadcast"); RT_SYM(AllReduce, "ncclAllReduce");
#undef RT_SYM
        api.ok = true;
    });
    return &api;
}

#define CKN(call)                                                                                         \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// rt_comm
// ---------------------------------------------------------------------------------------------
struct LocalGroup {            // shared by the rt_comm handles of one rt_comm_create_local call
    int n = 0;
    std::vector<rt_comm *> members;
    bool peer_ok = false;       // every device can map the root's (member 0's) memory and the root can map everyone's
    int refs = 0;
    std::mutex mu;
};

struct rt_comm {
    int n = 1, rank = 0, device = 0;
    ncclComm_t nccl = nullptr;
    LocalGroup *group = nullptr;
    float4 *frame = nullptr; size_t frame_cap = 0;          // W*H float4, grow-only
    uint8_t *rgba8 = nullptr; size_t rgba8_cap = 0;
    unsigned long long *d_cnt = nullptr;                    // 3 counters staged for their reduce
    // multi-process peer memory (CUDA IPC): the root's frame mapped into this rank's address space
    unsigned char *d_ipc = nullptr;                         // 64-byte cudaIpcMemHandle_t + 8 bytes of status, staged for the broadcast
    float4 *ipc_own = nullptr; size_t ipc_own_cap = 0;      // root: the exported frame (its own allocation: never reallocated behind the peers' backs)
    float4 *ipc_frame = nullptr;                            // root's exported frame as this rank sees it (root: == ipc_own)
    float4 *last_frame = nullptr;                           // where the last combined frame lives (rt_comm_frame)
    size_t ipc_px = 0; int ipc_root = -1;                   // what the mapping covers
    bool ipc_off = false;                                   // the ranks agreed that IPC does not work here (threads of one process, different nodes, ...)
    std::vector<uint32_t> tile_ids; uint32_t tile_w = 0, tile_h = 0, tile_sz = 0;     // cached tile partition ...
    uint32_t *d_tile_ids = nullptr; size_t d_tile_cap = 0;                            // ... and its device copy (uploaded once, not per frame)
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    PinnedHost pinned_out;                                  // the caller's float frame (RT_FLAG_PIN_HOST)
    double stats[4] = {0, 0, 0, 0};
};

static int comm_alloc(rt_comm **out, int n, int rank, int device) {
    CK(cudaSetDevice(device));
    rt_comm *c = new rt_comm;
    c->n = n; c->rank = rank; c->device = device;
    cudaError_t e = cudaMalloc((void **)&c->d_cnt, 3 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_ipc, 128);
    if (getenv("RT_B200_NO_IPC")) c->ipc_off = true;
    if (e == cudaSuccess) e = cudaEventCreate(&c->e0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->e1);
    if (e == cudaSuccess) e = cudaEventCreate(&c->e2);
    if (e != cudaSuccess) { rt_comm_destroy(c); return fail(RT_ERR_CUDA, "rt_comm allocation failed: %s", cudaGetErrorString(e)); }
    *out = c;
    return RT_OK;
}

extern "C" int rt_comm_unique_id(uint8_t out_id[RT_COMM_ID_BYTES]) {
    g_err.clear();
    if (!out_id) return fail(RT_ERR_ARG, "null argument");
    NcclApi *N = nccl_api();
    if (!N->ok) return fail(RT_ERR_STATE, "NCCL unavailable: %s", N->why.c_str());
    static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    CKN(N->GetUniqueId(&id));
    memcpy(out_id, &id, RT_COMM_ID_BYTES);
    return RT_OK;
}

extern "C" int rt_comm_create(int n_ranks, int rank, const uint8_t id[RT_COMM_ID_BYTES], int device, rt_comm **out_comm) {
    g_err.clear();
    if (!out_comm || (n_ranks > 1 && !id)) return fail(RT_ERR_ARG, "null argument");
    *out_comm = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(RT_ERR_ARG, "rank %d of %d", rank, n_ranks);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(RT_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    rt_comm *c = nullptr;
    int rc = comm_alloc(&c, n_ranks, rank, device);
    if (rc) return rc;
    if (n_ranks > 1) {
        NcclApi *N = nccl_api();
        if (!N->ok) { rt_comm_destroy(c); return fail(RT_ERR_STATE, "NCCL unavailable: %s", N->why.c_str()); }
        ncclUniqueId uid;
        memcpy(&uid, id, RT_COMM_ID_BYTES);
        ncclResult_t r = N->CommInitRank(&c->nccl, n_ranks, uid, rank);
        if (r != ncclSuccess) { rt_comm_destroy(c); return fail(RT_ERR_CUDA, "ncclCommInitRank failed: %s", N->GetErrorString(r)); }
    }
    *out_comm = c;
    return RT_OK;
}

extern "C" int rt_comm_create_local(int n, const int *devices, rt_comm **out_comms) {
    g_err.clear();
    if (!out_comms || n < 1) return fail(RT_ERR_ARG, "bad argument");
    for (int i = 0; i < n; ++i) out_comms[i] = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RT_ERR_CUDA, "no CUDA device: librt_b200 has no CPU fallback");
    std::vector<int> dev(n);
    for (int i = 0; i < n; ++i) {
        dev[i] = devices ? devices[i] : i;
        if (dev[i] < 0 || dev[i] >= ndev) return fail(RT_ERR_ARG, "device %d out of range (%d devices)", dev[i], ndev);
        for (int j = 0; j < i; ++j) if (dev[j] == dev[i]) return fail(RT_ERR_ARG, "device %d listed twice", dev[i]);
    }
    LocalGroup *g = new LocalGroup;
    g->n = n; g->refs = n; g->members.resize(n, nullptr);
    auto bail = [&](int rc) { for (int i = 0; i < n; ++i) if (out_comms[i]) { out_comms[i]->group = nullptr; rt_comm_destroy(out_comms[i]); out_comms[i] = nullptr; } delete g; return rc; };
    for (int i = 0; i < n; ++i) {
        int rc = comm_alloc(&out_comms[i], n, i, dev[i]);
        if (rc) return bail(rc);
        out_comms[i]->group = g; g->members[i] = out_comms[i];
    }
    // peer access: everyone <-> root (device of rank 0)
    bool peer = n > 1;
    for (int i = 1; i < n && peer; ++i) {
        int a = 0, b = 0;
        if (cudaDeviceCanAccessPeer(&a, dev[i], dev[0]) != cudaSuccess || cudaDeviceCanAccessPeer(&b, dev[0], dev[i]) != cudaSuccess || !a || !b) peer = false;
    }
    if (peer && getenv("RT_B200_NO_P2P")) peer = false;
    if (peer) {
        for (int i = 1; i < n && peer; ++i) {
            cudaError_t e;
            cudaSetDevice(dev[i]); e = cudaDeviceEnablePeerAccess(dev[0], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer = false;
            cudaSetDevice(dev[0]); e = cudaDeviceEnablePeerAccess(dev[i], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer = false;
            (void)cudaGetLastError();
        }
    }
    g->peer_ok = peer;
    if (!peer && n > 1) {          // NCCL fallback: one communicator per device
        NcclApi *N = nccl_api();
        if (!N->ok) return bail(fail(RT_ERR_STATE, "no peer access between the GPUs and NCCL unavailable: %s", N->why.c_str()));
        std::vector<ncclComm_t> comms(n);
        ncclResult_t r = N->CommInitAll(comms.data(), n, dev.data());
        if (r != ncclSuccess) return bail(fail(RT_ERR_CUDA, "ncclCommInitAll failed: %s", N->GetErrorString(r)));
        for (int i = 0; i < n; ++i) out_comms[i]->nccl = comms[i];
    }
    return RT_OK;
}

extern "C" void rt_comm_destroy(rt_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    c->pinned_out.release();
    if (c->nccl && nccl_api()->ok) nccl_api()->CommDestroy(c->nccl);
    if (c->ipc_frame && c->ipc_frame != c->ipc_own) cudaIpcCloseMemHandle(c->ipc_frame);
    if (c->ipc_own) cudaFree(c->ipc_own);
    if (c->frame) cudaFree(c->frame);
    if (c->rgba8) cudaFree(c->rgba8);
    if (c->d_ipc) cudaFree(c->d_ipc);
    if (c->d_tile_ids) cudaFree(c->d_tile_ids);
    if (c->d_cnt) cudaFree(c->d_cnt);
    if (c->e0) cudaEventDestroy(c->e0);
    if (c->e1) cudaEventDestroy(c->e1);
    if (c->e2) cudaEventDestroy(c->e2);
    if (c->group) {
        LocalGroup *g = c->group;
        bool last;
        { std::lock_guard<std::mutex> lk(g->mu); g->members[c->rank] = nullptr; last = --g->refs == 0; }
        if (last) delete g;
    }
    delete c;
}

extern "C" int rt_comm_rank(const rt_comm *c) { return c ? c->rank : -1; }
extern "C" int rt_comm_size(const rt_comm *c) { return c ? c->n : 0; }
extern "C" const float *rt_comm_frame(const rt_comm *c) { return c ? (const float *)c->last_frame : nullptr; }
extern "C" int rt_comm_get_stats(const rt_comm *c, double out[4]) {
    if (!c || !out) return fail(RT_ERR_ARG, "null argument");
    memcpy(out, c->stats, sizeof(c->stats));
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// partitions (host)
// ---------------------------------------------------------------------------------------------
extern "C" int rt_partition_tiles(uint32_t width, uint32_t height, uint32_t tile, int rank, int n_ranks, uint32_t *out_ids, uint32_t *out_count) {
    g_err.clear();
    if (!out_count) return fail(RT_ERR_ARG, "null argument");
    if (!tile || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(RT_ERR_ARG, "bad tile partition (tile %u, rank %d of %d)", tile, rank, n_ranks);
    const uint32_t tx_n = (width + tile - 1) / tile, ty_n = (height + tile - 1) / tile;
    uint64_t count = 0;
    for (uint64_t t = (uint64_t)rank; t < (uint64_t)tx_n * ty_n; t += (uint64_t)n_ranks) {
        const uint32_t ty = (uint32_t)(t / tx_n), tx = (uint32_t)(t % tx_n);
        const uint32_t x0 = tx * tile, y0 = ty * tile, x1 = std::min(x0 + tile, width), y1 = std::min(y0 + tile, height);
        if (out_ids) for (uint32_t y = y0; y < y1; ++y) for (uint32_t x = x0; x < x1; ++x) out_ids[count++] = y * width + x;
        else count += (uint64_t)(x1 - x0) * (y1 - y0);
    }
    *out_count = (uint32_t)count;
    return RT_OK;
}

static void sample_partition(uint32_t total, int rank, int n, uint32_t *begin, uint32_t *count) {      // remainders go to the lowest ranks
    const uint32_t base = total / (uint32_t)n, rem = total % (uint32_t)n;
    *begin = (uint32_t)rank * base + std::min((uint32_t)rank, rem);
    *count = base + ((uint32_t)rank < rem ? 1u : 0u);
}

static void range_partition(uint64_t total, int rank, int n, uint32_t *begin, uint32_t *count) {       // main.cpp:311-317, clipped to the frame
    const uint64_t cpp = (total + (uint64_t)n - 1) / (uint64_t)n;
    const uint64_t a = std::min(total, cpp * (uint64_t)rank), b = std::min(total, cpp * ((uint64_t)rank + 1));
    *begin = (uint32_t)a; *count = (uint32_t)(b - a);
}

// ---------------------------------------------------------------------------------------------
// resolve kernels of the combine
// ---------------------------------------------------------------------------------------------
// sample partition, after the reduce: frame holds the sum over ALL samples in xyz -> color /= samp; color.w = 1 (main.cpp:262-263)
__global__ void k_samples_resolve(float4 *frame, uint32_t n, float total) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 c = frame[p];
    frame[p] = make_float4(c.x / total, c.y / total, c.z / total, 1.0f);
}

// sample partition through peer memory: the root adds the ranks' sums in RANK ORDER (a fixed order, so the frame does not depend on
// the interconnect's reduction tree) reading its peers' frames over NVLink, and resolves in the same pass.
#define RT_MAX_LOCAL 16
struct PeerFrames { const float4 *f[RT_MAX_LOCAL]; };
__global__ void k_peer_sum_resolve(PeerFrames P, int n_ranks, float4 *out, uint32_t n, float total) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 c = P.f[0][p];
    for (int r = 1; r < n_ranks; ++r) { float4 a = P.f[r][p]; c.x += a.x; c.y += a.y; c.z += a.z; }
    out[p] = make_float4(c.x / total, c.y / total, c.z / total, 1.0f);
}

// ---------------------------------------------------------------------------------------------
// rt_render_combined
// ---------------------------------------------------------------------------------------------
template <typename T> static int grow_dev(T **buf, size_t *cap, size_t need) {
    if (*buf && *cap >= need) return RT_OK;
    if (*buf) CK(cudaFree(*buf));
    *buf = nullptr; *cap = 0;
    CK(cudaMalloc((void **)buf, std::max<size_t>(need, 1) * sizeof(T)));
    *cap = need;
    return RT_OK;
}

static int check_combined_args(rt_scene *scene, rt_comm *comm, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                               int partition, uint32_t tile, uint32_t flags) {
    if (!scene || !comm || !cam || !params) return fail(RT_ERR_ARG, "null argument");
    if (scene->device != comm->device) return fail(RT_ERR_ARG, "scene lives on device %d, comm on device %d", scene->device, comm->device);
    if (!width || !height || (uint64_t)width * height > 0xFFFFFFFFull) return fail(RT_ERR_ARG, "bad frame size");
    if (partition != RT_PART_TILES && partition != RT_PART_RANGES && partition != RT_PART_SAMPLES) return fail(RT_ERR_ARG, "unknown partition %d", partition);
    if (partition == RT_PART_TILES && !tile) return fail(RT_ERR_ARG, "tile size 0");
    if (partition == RT_PART_SAMPLES && (flags & RT_FLAG_ADAPTIVE) && params->min_samples < params->max_samples)
        return fail(RT_ERR_ARG, "adaptive sampling decides per pixel: use a pixel partition");
    if (flags & ~(uint32_t)(RT_FLAG_ADAPTIVE | RT_FLAG_TIME_KERNELS | RT_FLAG_COUNTERS | RT_FLAG_PIN_HOST)) return fail(RT_ERR_ARG, "flags 0x%x not accepted here", flags);
    return RT_OK;
}

// this rank's share of the frame, rendered into `dst` (a W*H frame: its own, or -- peer mode, pixel partitions -- the root's). The render
// is ordered after the device's legacy default stream (stream NULL), like every rt_render_device call without a stream.
static int render_share(rt_scene *scene, rt_comm *comm, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                        int partition, uint32_t tile, uint32_t flags, float4 *dst, rt_counters *cnt) {
    const uint32_t n_px = width * height;
    if (partition == RT_PART_SAMPLES) {
        uint32_t s0, ns;
        sample_partition(params->min_samples, comm->rank, comm->n, &s0, &ns);
        return rt_render_device(scene, cam, params, width, height, nullptr, 0, n_px, s0, ns, flags | RT_OUT_SUM | RT_OUT_FULLFRAME, (float *)dst, nullptr, cnt);
    }
    if (partition == RT_PART_RANGES) {
        uint32_t p0, np;
        range_partition(n_px, comm->rank, comm->n, &p0, &np);
        return rt_render_device(scene, cam, params, width, height, nullptr, p0, np, 0, params->min_samples, flags | RT_OUT_MEAN | RT_OUT_FULLFRAME, (float *)dst, nullptr, cnt);
    }
    if (comm->tile_w != width || comm->tile_h != height || comm->tile_sz != tile) {
        uint32_t count = 0;
        int rc = rt_partition_tiles(width, height, tile, comm->rank, comm->n, nullptr, &count);
        if (rc) return rc;
        comm->tile_ids.resize(count);
        rc = rt_partition_tiles(width, height, tile, comm->rank, comm->n, comm->tile_ids.data(), &count);
        if (rc) return rc;
        rc = grow_dev(&comm->d_tile_ids, &comm->d_tile_cap, (size_t)count);
        if (rc) return rc;
        CK(cudaMemcpy(comm->d_tile_ids, comm->tile_ids.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
        comm->tile_w = width; comm->tile_h = height; comm->tile_sz = tile;
    }
    return rt_render_device_ids(scene, cam, params, width, height, comm->d_tile_ids, (uint32_t)comm->tile_ids.size(), 0, params->min_samples,
                                flags | RT_OUT_MEAN | RT_OUT_FULLFRAME, (float *)dst, cnt);
}

// root only: optional tone map + downloads of the finished frame
static int deliver(rt_scene *scene, rt_comm *comm, const float4 *frame, uint32_t width, uint32_t height, float *out_rgba_host, uint8_t *out_rgba8_host,
                   float *out_scene_luma) {
    const size_t n_px = (size_t)width * height;
    cudaStream_t st = scene->stream;
    CK(cudaEventRecord(comm->e1, st));
    if (out_rgba8_host || out_scene_luma) {
        int rc = grow_dev(&comm->rgba8, &comm->rgba8_cap, n_px * 4);
        if (rc) return rc;
        rc = rt_tonemap_device(comm->device, (const float *)frame, width, height, comm->rgba8, out_scene_luma, st);
        if (rc) return rc;
        if (out_rgba8_host) { CK(cudaMemcpyAsync(out_rgba8_host, comm->rgba8, n_px * 4, cudaMemcpyDeviceToHost, st)); scene->stats.d2h_bytes += n_px * 4; }
    }
    if (out_rgba_host) { CK(cudaMemcpyAsync(out_rgba_host, frame, n_px * 16, cudaMemcpyDeviceToHost, st)); scene->stats.d2h_bytes += n_px * 16; }
    CK(cudaEventRecord(comm->e2, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, comm->e1, comm->e2));
    comm->stats[1] = ms;
    return RT_OK;
}

// Multi-process peer memory: every rank maps the ROOT's frame (CUDA IPC) and its resolve kernel stores its pixels straight into it over
// NVLink -- the pixel partitions cover the frame, so there is neither a zero-fill nor a reduce of full frames. Collective; returns with
// comm->ipc_frame set on every rank, or with comm->ipc_off set on every rank (the ranks agree through an all-reduce) when the mapping is
// not possible (ranks that are threads of one process, ranks on different nodes, IPC disabled): then the NCCL reduce is used.
static int ensure_ipc(rt_comm *comm, int root, size_t n_px, cudaStream_t st) {
    if (comm->ipc_off) return RT_OK;
    if (comm->ipc_frame && comm->ipc_root == root && comm->ipc_px >= n_px) return RT_OK;      // same decision on every rank: same arguments
    NcclApi *N = nccl_api();
    const bool is_root = comm->rank == root;
    if (comm->ipc_frame && comm->ipc_frame != comm->ipc_own) cudaIpcCloseMemHandle(comm->ipc_frame);     // before the root frees what it maps
    comm->ipc_frame = nullptr; comm->ipc_px = 0; comm->ipc_root = -1;
    int *d_flag = reinterpret_cast<int *>(comm->d_ipc + 80);
    {   // barrier: the root may reallocate its exported frame only after every rank has closed its mapping of the old one
        int one = 1;
        CK(cudaMemcpyAsync(d_flag, &one, sizeof(int), cudaMemcpyHostToDevice, st));
        CKN(N->AllReduce(d_flag, d_flag, 1, ncclInt32, ncclMin, comm->nccl, st));
        CK(cudaStreamSynchronize(st));
    }
    unsigned char hbuf[72];
    memset(hbuf, 0, sizeof(hbuf));
    if (is_root) {
        int rc = grow_dev(&comm->ipc_own, &comm->ipc_own_cap, n_px);
        if (rc) return rc;
        cudaIpcMemHandle_t h;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
        const cudaError_t e = cudaIpcGetMemHandle(&h, comm->ipc_own);
        if (e == cudaSuccess) { memcpy(hbuf, &h, 64); hbuf[64] = 1; } else (void)cudaGetLastError();
        CK(cudaMemcpyAsync(comm->d_ipc, hbuf, 72, cudaMemcpyHostToDevice, st));
    }
    CKN(N->Broadcast(comm->d_ipc, comm->d_ipc, 72, ncclUint8, root, comm->nccl, st));
    CK(cudaMemcpyAsync(hbuf, comm->d_ipc, 72, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int ok = hbuf[64] == 1;
    float4 *mapped = nullptr;
    if (ok) {
        if (is_root) mapped = comm->ipc_own;
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, hbuf, 64);
            void *ptr = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { ok = 0; (void)cudaGetLastError(); } else mapped = (float4 *)ptr;
        }
    }
    // agree: all or nothing
    CK(cudaMemcpyAsync(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
    CKN(N->AllReduce(d_flag, d_flag, 1, ncclInt32, ncclMin, comm->nccl, st));
    int all = 0;
    CK(cudaMemcpyAsync(&all, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (!all) {
        if (mapped && !is_root) cudaIpcCloseMemHandle(mapped);
        comm->ipc_off = true;
        return RT_OK;
    }
    comm->ipc_frame = mapped; comm->ipc_px = n_px; comm->ipc_root = root;
    return RT_OK;
}

extern "C" int rt_render_combined(rt_scene *scene, rt_comm *comm, const rt_camera *cam, const rt_params *params, uint32_t width, uint32_t height,
                                  int partition, uint32_t tile, uint32_t flags, int root, float *out_rgba_host, uint8_t *out_rgba8_host,
                                  float *out_scene_luma, rt_counters *out_counters) {
    g_err.clear();
    int rc = check_combined_args(scene, comm, cam, params, width, height, partition, tile, flags);
    if (rc) return rc;
    if (root < 0 || root >= comm->n) return fail(RT_ERR_ARG, "root %d of %d ranks", root, comm->n);
    if (comm->n > 1 && !comm->nccl) return fail(RT_ERR_STATE, "this rt_comm has no NCCL communicator (peer-memory group: use rt_render_multi)");
    CK(cudaSetDevice(comm->device));
    if ((flags & RT_FLAG_PIN_HOST) && out_rgba_host && comm->rank == root) comm->pinned_out.pin(out_rgba_host, (size_t)width * height * 16);
    flags &= ~(uint32_t)RT_FLAG_PIN_HOST;
    const uint32_t n_px = width * height;
    cudaStream_t st = scene->stream;
    const bool is_root = comm->rank == root;
    memset(comm->stats, 0, sizeof(comm->stats));

    // Pixel partitions between processes that share NVLink: peer-memory stores into the root's frame (ensure_ipc). One tiny broadcast
    // from the root FIRST: the root enqueues it only after it has consumed (downloaded / tone-mapped) the previous frame, and every other
    // rank's stores of this frame are ordered after it on its stream -- nobody overwrites a frame the root is still reading.
    bool ipc = false;
    if (comm->n > 1 && partition != RT_PART_SAMPLES && !comm->ipc_off) {
        rc = ensure_ipc(comm, root, (size_t)n_px, st);
        if (rc) return rc;
        ipc = !comm->ipc_off && comm->ipc_frame != nullptr;
        if (ipc) CKN(nccl_api()->Broadcast(comm->d_ipc + 96, comm->d_ipc + 96, 4, ncclUint8, root, comm->nccl, st));
    }
    if (!ipc) { rc = grow_dev(&comm->frame, &comm->frame_cap, (size_t)n_px); if (rc) return rc; }
    float4 *const frame = ipc ? comm->ipc_frame : comm->frame;      // where this rank's pixels go; on the root: the combined frame
    comm->last_frame = frame;
    // NCCL reduce of pixel partitions: the other ranks' pixels must read as 0 for the sum to equal the gather
    if (!ipc && (comm->n > 1 || partition != RT_PART_SAMPLES)) CK(cudaMemsetAsync(frame, 0, (size_t)n_px * sizeof(float4), st));
    rt_counters cnt;
    memset(&cnt, 0, sizeof(cnt));
    rc = render_share(scene, comm, cam, params, width, height, partition, tile, flags, frame, &cnt);
    if (rc) return rc;

    CK(cudaEventRecord(comm->e0, st));
    if (comm->n > 1) {
        NcclApi *N = nccl_api();
        if (!ipc) {
            CKN(N->Reduce(frame, frame, (size_t)n_px * 4, ncclFloat, ncclSum, root, comm->nccl, st));
            comm->stats[2] = (double)n_px * 16.0;
        } else comm->stats[3] = 1.0;
        // the counters' reduce doubles as the completion signal of the peer-memory gather: it finishes on the root only after every rank has
        // enqueued it, i.e. after every rank's resolve kernel (and its stores into the root's frame) has completed
        unsigned long long hc[3] = {cnt.ray_count, cnt.sphere_check_count, cnt.mesh_check_count};
        CK(cudaMemcpyAsync(comm->d_cnt, hc, sizeof(hc), cudaMemcpyHostToDevice, st));
        CKN(N->Reduce(comm->d_cnt, comm->d_cnt, 3, ncclUint64, ncclSum, root, comm->nccl, st));
        if (is_root) {
            CK(cudaMemcpyAsync(hc, comm->d_cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            cnt.ray_count = hc[0]; cnt.sphere_check_count = hc[1]; cnt.mesh_check_count = hc[2];
        }
    }
    if (partition == RT_PART_SAMPLES && is_root) {
        k_samples_resolve<<<cdiv(n_px, 256), 256, 0, st>>>(frame, n_px, (float)params->min_samples);
        CKL("k_samples_resolve");
        scene->stats.kernel_launches += 1;
    }
    CK(cudaEventRecord(comm->e1, st));
    if (is_root && (out_rgba_host || out_rgba8_host || out_scene_luma)) {
        rc = deliver(scene, comm, frame, width, height, out_rgba_host, out_rgba8_host, out_scene_luma);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(st));
    { float ms = 0; CK(cudaEventElapsedTime(&ms, comm->e0, comm->e1)); comm->stats[0] = ms; }
    if (out_counters) *out_counters = cnt;
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// rt_render_multi: one process, n GPUs, one host thread per GPU
// ---------------------------------------------------------------------------------------------
extern "C" int rt_render_multi(rt_scene *const *scenes, rt_comm *const *comms, int n, const rt_camera *cam, const rt_params *params,
                               uint32_t width, uint32_t height, int partition, uint32_t tile, uint32_t flags,
                               float *out_rgba_host, uint8_t *out_rgba8_host, float *out_scene_luma, rt_counters *out_counters) {
    g_err.clear();
    if (!scenes || !comms || n < 1 || n > RT_MAX_LOCAL) return fail(RT_ERR_ARG, "bad argument (1 <= n <= %d)", RT_MAX_LOCAL);
    for (int i = 0; i < n; ++i) {
        if (!scenes[i] || !comms[i]) return fail(RT_ERR_ARG, "null scene / comm %d", i);
        int rc = check_combined_args(scenes[i], comms[i], cam, params, width, height, partition, tile, flags);
        if (rc) return rc;
        if (comms[i]->n != n || comms[i]->rank != i || !comms[i]->group || comms[i]->group != comms[0]->group)
            return fail(RT_ERR_ARG, "comms[%d] is not rank %d of one rt_comm_create_local group of %d", i, i, n);
    }
    if ((flags & RT_FLAG_PIN_HOST) && out_rgba_host) { CK(cudaSetDevice(comms[0]->device)); comms[0]->pinned_out.pin(out_rgba_host, (size_t)width * height * 16); }
    flags &= ~(uint32_t)RT_FLAG_PIN_HOST;
    LocalGroup *g = comms[0]->group;
    const bool peer = g->peer_ok || n == 1;
    const uint32_t n_px = width * height;
    std::vector<int> rcs(n, RT_OK);
    std::vector<std::string> errs(n);
    std::vector<rt_counters> cnts(n);
    memset(cnts.data(), 0, sizeof(rt_counters) * n);

    if (!peer) {       // NCCL between the threads: each runs the whole combined call; only the root delivers
        std::vector<std::thread> th;
        for (int i = 0; i < n; ++i)
            th.emplace_back([&, i] {
                rcs[i] = rt_render_combined(scenes[i], comms[i], cam, params, width, height, partition, tile, flags, 0,
                                            i == 0 ? out_rgba_host : nullptr, i == 0 ? out_rgba8_host : nullptr, i == 0 ? out_scene_luma : nullptr, &cnts[i]);
                if (rcs[i]) errs[i] = rt_last_error();
            });
        for (auto &t : th) t.join();
        for (int i = 0; i < n; ++i) if (rcs[i]) return fail(rcs[i], "GPU %d: %s", comms[i]->device, errs[i].c_str());
        if (out_counters) *out_counters = cnts[0];
        return RT_OK;
    }

    // ---- peer memory ----
    rt_comm *root = comms[0];
    CK(cudaSetDevice(root->device));
    { int rc = grow_dev(&root->frame, &root->frame_cap, (size_t)n_px); if (rc) return rc; }
    memset(root->stats, 0, sizeof(root->stats));
    root->last_frame = root->frame;
    const bool samples = partition == RT_PART_SAMPLES;
    if (samples) for (int i = 1; i < n; ++i) { CK(cudaSetDevice(comms[i]->device)); int rc = grow_dev(&comms[i]->frame, &comms[i]->frame_cap, (size_t)n_px); if (rc) return rc; }
    {
        // Pixel partitions: every GPU's k_finalize stores its pixels straight into the ROOT's frame (disjoint pixels, together they cover
        // the frame: no zero-fill, no reduce). rt_render_device returns after its stream has drained, so the stores are complete at join.
        std::vector<std::thread> th;
        for (int i = 0; i < n; ++i)
            th.emplace_back([&, i] {
                if (cudaSetDevice(comms[i]->device) != cudaSuccess) { rcs[i] = RT_ERR_CUDA; errs[i] = "cudaSetDevice failed"; return; }
                rcs[i] = render_share(scenes[i], comms[i], cam, params, width, height, partition, tile, flags, samples ? comms[i]->frame : root->frame, &cnts[i]);
                if (rcs[i]) errs[i] = rt_last_error();
            });
        for (auto &t : th) t.join();
        for (int i = 0; i < n; ++i) if (rcs[i]) return fail(rcs[i], "GPU %d: %s", comms[i]->device, errs[i].c_str());
    }
    CK(cudaSetDevice(root->device));
    cudaStream_t st = scenes[0]->stream;
    CK(cudaEventRecord(root->e0, st));
    if (samples) {
        PeerFrames P;
        for (int i = 0; i < n; ++i) P.f[i] = comms[i]->frame;
        k_peer_sum_resolve<<<cdiv(n_px, 256), 256, 0, st>>>(P, n, root->frame, n_px, (float)params->min_samples);
        CKL("k_peer_sum_resolve");
        scenes[0]->stats.kernel_launches += 1;
        root->stats[2] = (double)n_px * 16.0;
    }
    CK(cudaEventRecord(root->e1, st));
    root->stats[3] = 1.0;
    if (out_rgba_host || out_rgba8_host || out_scene_luma) {
        int rc = deliver(scenes[0], root, root->frame, width, height, out_rgba_host, out_rgba8_host, out_scene_luma);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(st));
    { float ms = 0; CK(cudaEventElapsedTime(&ms, root->e0, root->e1)); root->stats[0] = ms; }
    if (out_counters) {
        memset(out_counters, 0, sizeof(*out_counters));
        for (int i = 0; i < n; ++i) { out_counters->ray_count += cnts[i].ray_count; out_counters->sphere_check_count += cnts[i].sphere_check_count; out_counters->mesh_check_count += cnts[i].mesh_check_count; }
    }
    return RT_OK;
}
