// store_peak.cu — what pure-store kernels reach on this GPU (the ceiling of a Matte8 resolve, which
// writes 1 B/px and reads nothing).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_peak store_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// (a) grid-stride 16-byte stores
__global__ void k_stride(uint4 *p, size_t n, uint32_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(v, v, v, v);
}
// (b) the tile kernel's pattern: a warp owns a tile of `rows` rows of `row_bytes` bytes and writes it row by row
__global__ void k_tiles(uint4 *p, size_t n_tiles, uint32_t rows, uint32_t row_u4, uint32_t v) {
    const uint32_t lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const size_t n_warps = (size_t)gridDim.x * wpc;
    for (size_t t = blockIdx.x * (size_t)wpc + (threadIdx.x >> 5); t < n_tiles; t += n_warps) {
        uint4 *q = p + t * rows * row_u4;
        for (uint32_t r = 0; r < rows; r++, q += row_u4)
            for (uint32_t g = lane; g < row_u4; g += 32) q[g] = make_uint4(v, v, v, v);
    }
}
int main() {
    const size_t bytes = 256ull * 4096 * 4096;  // the bench step: 256 rasters of 4096^2
    uint4 *d;
    CK(cudaMalloc(&d, bytes));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int it = 0; it < 3; it++) CK(cudaMemset(d, 1, bytes));
    cudaEventRecord(a); for (int it = 0; it < 10; it++) cudaMemset(d, it, bytes); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&ms, a, b); printf("cudaMemset           %.3f ms  %.1f GB/s\n", ms / 10, bytes / (ms / 10) / 1e6);
    for (int ctas = 4; ctas <= 16; ctas *= 2) {
        for (int it = 0; it < 3; it++) k_stride<<<148 * ctas, 128>>>(d, bytes / 16, it);
        cudaEventRecord(a); for (int it = 0; it < 10; it++) k_stride<<<148 * ctas, 128>>>(d, bytes / 16, it); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b); printf("grid-stride %2d CTA/SM %.3f ms  %.1f GB/s\n", ctas, ms / 10, bytes / (ms / 10) / 1e6);
    }
    for (int ctas = 4; ctas <= 8; ctas += 2) for (uint32_t rows = 4; rows <= 16; rows *= 2) {
        const size_t n_tiles = bytes / (4096ull * rows);
        for (int it = 0; it < 3; it++) k_tiles<<<148 * ctas, 128>>>(d, n_tiles, rows, 256, it);
        cudaEventRecord(a); for (int it = 0; it < 10; it++) k_tiles<<<148 * ctas, 128>>>(d, n_tiles, rows, 256, it); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b); printf("warp tiles %2d CTA/SM %2u rows %.3f ms  %.1f GB/s\n", ctas, rows, ms / 10, bytes / (ms / 10) / 1e6);
    }
    return 0;
}
