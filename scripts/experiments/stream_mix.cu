// Microbenchmark: what HBM bandwidth does k_logic's access MIX reach when nothing else is in the way?
// Each thread handles one path slot: reads NR float4 streams, writes NW float4 streams (SoA, 16 B per thread per stream), persistent
// grid like k_logic (1776 blocks x 128). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_mix stream_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NR, int NW>
__global__ void __launch_bounds__(128, 6) k_mix(float4 *const *rd, float4 *const *wr, unsigned n) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < NR; ++k) { float4 v = rd[k][i]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
#pragma unroll
        for (int k = 0; k < NW; ++k) { float4 o = a; o.x += k; wr[k][i] = o; }
    }
}
template <int NR, int NW> void run(unsigned n) {
    float4 *h_rd[16], *h_wr[16], **d_rd, **d_wr;
    for (int k = 0; k < NR; ++k) { cudaMalloc(&h_rd[k], (size_t)n * 16); cudaMemset(h_rd[k], 0, (size_t)n * 16); }
    for (int k = 0; k < NW; ++k) cudaMalloc(&h_wr[k], (size_t)n * 16);
    cudaMalloc(&d_rd, sizeof(h_rd)); cudaMalloc(&d_wr, sizeof(h_wr));
    cudaMemcpy(d_rd, h_rd, sizeof(h_rd), cudaMemcpyHostToDevice); cudaMemcpy(d_wr, h_wr, sizeof(h_wr), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_mix<NR, NW><<<1776, 128>>>(d_rd, d_wr, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("reads %2d writes %2d streams, %u slots: %.3f ms, %.2f TB/s\n", NR, NW, n, ms, (double)n * 16 * (NR + NW) / ms / 1e9);
    }
    for (int k = 0; k < NR; ++k) cudaFree(h_rd[k]);
    for (int k = 0; k < NW; ++k) cudaFree(h_wr[k]);
    cudaFree(d_rd); cudaFree(d_wr);
}
int main() {
    const unsigned n = 1u << 25;
    run<1, 1>(n); run<4, 4>(n); run<2, 12>(n); run<8, 8>(n); run<12, 2>(n); run<1, 13>(n);
    return 0;
}
