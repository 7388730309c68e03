// rt_sort.cuh -- bitonic sort of (u64 key, u32 value) pairs and a packed-counter prefix sum. `static` kernels: included by the scene
// build (rt_scene.cu) and by the load-time tangent pass (rt_loadtime.cu), each translation unit gets its own copy.
#pragma once
#include "rt_common.cuh"

// ---- 2. bitonic sort of (key, val), lexicographic so the order is total and deterministic --------

RT_DEVICE bool kv_greater(uint64_t ka, uint32_t va, uint64_t kb, uint32_t vb) { return ka > kb || (ka == kb && va > vb); }

static __global__ void k_bitonic_global(uint64_t *keys, uint32_t *vals, uint32_t n_pad, uint32_t j, uint32_t k) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    uint32_t l = i ^ j;
    if (l > i) {
        uint64_t ka = keys[i], kb = keys[l];
        uint32_t va = vals[i], vb = vals[l];
        bool up = (i & k) == 0;
        if (kv_greater(ka, va, kb, vb) == up) { keys[i] = kb; keys[l] = ka; vals[i] = vb; vals[l] = va; }
    }
}

#define BITONIC_TILE 2048
// all (k, j) steps with j < BITONIC_TILE for k in [k_begin, k_end] (powers of two), inside shared memory
static __global__ void __launch_bounds__(1024) k_bitonic_shared(uint64_t *keys, uint32_t *vals, uint32_t k_begin, uint32_t k_end, int only_tail) {
    __shared__ uint64_t sk[BITONIC_TILE];
    __shared__ uint32_t sv[BITONIC_TILE];
    uint32_t base = blockIdx.x * BITONIC_TILE;
    for (uint32_t t = threadIdx.x; t < BITONIC_TILE; t += blockDim.x) { sk[t] = keys[base + t]; sv[t] = vals[base + t]; }
    __syncthreads();
    for (uint32_t k = k_begin; k <= k_end; k <<= 1) {
        uint32_t j0 = only_tail ? (BITONIC_TILE >> 1) : (k >> 1);
        if (j0 > (BITONIC_TILE >> 1)) j0 = BITONIC_TILE >> 1;
        for (uint32_t j = j0; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < BITONIC_TILE; t += blockDim.x) {
                uint32_t l = t ^ j;
                if (l > t) {
                    bool up = ((base + t) & k) == 0;
                    if (kv_greater(sk[t], sv[t], sk[l], sv[l]) == up) {
                        uint64_t tk = sk[t]; sk[t] = sk[l]; sk[l] = tk;
                        uint32_t tv = sv[t]; sv[t] = sv[l]; sv[l] = tv;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t t = threadIdx.x; t < BITONIC_TILE; t += blockDim.x) { keys[base + t] = sk[t]; vals[base + t] = sv[t]; }
}

// ---- prefix sum over packed (valid, merge) counters ----------------------------------------------

#define SCAN_BLOCK 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)

static __global__ void __launch_bounds__(SCAN_BLOCK) k_scan_reduce(const uint64_t *in, uint32_t n, uint64_t *block_sums) {
    __shared__ uint64_t sh[SCAN_BLOCK / 32];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) s += in[base + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < SCAN_BLOCK / 32; ++w) t += sh[w]; block_sums[blockIdx.x] = t; }
}

static __global__ void __launch_bounds__(1024) k_scan_blocksums(uint64_t *block_sums, uint32_t nb, uint64_t *total) {
    __shared__ uint64_t sh[1024];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = i < nb ? block_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t o = 1; o < 1024; o <<= 1) {
            uint64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        uint64_t incl = sh[threadIdx.x];
        if (i < nb) block_sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

static __global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(const uint64_t *in, uint32_t n, const uint64_t *block_sums, uint64_t *out) {
    __shared__ uint64_t sh[SCAN_BLOCK];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t o = 1; o < SCAN_BLOCK; o <<= 1) {
        uint64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint64_t run = block_sums[blockIdx.x] + sh[threadIdx.x] - s;
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
}
