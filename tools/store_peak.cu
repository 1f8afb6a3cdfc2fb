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
// (c) the analytic rows' pattern: every row is 4 constant spans with a single-lane 16-byte "edge group"
// between them, at 16-byte-aligned positions that move from row to row (spans do not start on 128-byte lines)
__global__ void k_spans(uint4 *p, size_t n_tiles, uint32_t rows, uint32_t row_u4, uint32_t v, int line_aligned) {
    const uint32_t lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const size_t n_warps = (size_t)gridDim.x * wpc;
    for (size_t t = blockIdx.x * (size_t)wpc + (threadIdx.x >> 5); t < n_tiles; t += n_warps) {
        uint4 *q = p + t * rows * row_u4;
        for (uint32_t r = 0; r < rows; r++, q += row_u4) {
            uint32_t cut[5];
            cut[0] = 0;
            for (int k = 1; k < 4; k++) {
                cut[k] = (uint32_t)((t * 7 + r * 3 + k * 61) % 60) + k * 64 - 30;  // group index of the k-th edge
                if (line_aligned) cut[k] &= ~7u;
            }
            cut[4] = row_u4;
            for (int k = 0; k < 4; k++) {
                const uint32_t lo = cut[k] + (k ? 1 : 0), hi = cut[k + 1];
                if (k && lane == 0) q[cut[k]] = make_uint4(v + 1, v, v, v);  // the edge group, one lane
                for (uint32_t g = lo + lane; g < hi; g += 32) q[g] = make_uint4(v, v, v, v);
            }
        }
    }
}
// (d)/(e) as (c), but the constant spans cover whole 128-byte lines only and the line holding an edge is
// written by ONE lane as 8 x 16 bytes (mode 0) or by 8 lanes in one instruction (mode 1)
__global__ void k_lines(uint4 *p, size_t n_tiles, uint32_t rows, uint32_t row_u4, uint32_t v, int mode) {
    const uint32_t lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const size_t n_warps = (size_t)gridDim.x * wpc;
    for (size_t t = blockIdx.x * (size_t)wpc + (threadIdx.x >> 5); t < n_tiles; t += n_warps) {
        uint4 *q = p + t * rows * row_u4;
        for (uint32_t r = 0; r < rows; r++, q += row_u4) {
            uint32_t cut[5];
            cut[0] = 0;
            for (int k = 1; k < 4; k++) cut[k] = (uint32_t)((t * 7 + r * 3 + k * 61) % 60) + k * 64 - 30;
            cut[4] = row_u4;
            for (int k = 0; k < 4; k++) {
                const uint32_t lo = k ? (cut[k] | 7u) + 1u : 0u, hi = cut[k + 1] & ~7u;  // whole lines
                if (k) {
                    const uint32_t l0 = cut[k] & ~7u;
                    if (mode == 0) {
                        if (lane == 0)
                            for (uint32_t i = 0; i < 8; i++) q[l0 + i] = make_uint4(v + (l0 + i == cut[k]), v, v, v);
                    } else if (lane < 8) q[l0 + lane] = make_uint4(v + (l0 + lane == cut[k]), v, v, v);
                }
                for (uint32_t g = lo + lane; g < hi; g += 32) q[g] = make_uint4(v, v, v, v);
            }
        }
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
    for (int al = 0; al < 2; al++) {
        const uint32_t rows = 8; const int ctas = 5;
        const size_t n_tiles = bytes / (4096ull * rows);
        for (int it = 0; it < 3; it++) k_spans<<<148 * ctas, 128>>>(d, n_tiles, rows, 256, it, al);
        cudaEventRecord(a); for (int it = 0; it < 10; it++) k_spans<<<148 * ctas, 128>>>(d, n_tiles, rows, 256, it, al); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b); printf("warp spans + edge groups (%s) %.3f ms  %.1f GB/s\n", al ? "edges on 128-byte lines" : "edges anywhere", ms / 10, bytes / (ms / 10) / 1e6);
    }
    for (int mode = 0; mode < 2; mode++) {
        const uint32_t rows = 8; const int ctas = 5;
        const size_t n_tiles = bytes / (4096ull * rows);
        for (int it = 0; it < 3; it++) k_lines<<<148 * ctas, 128>>>(d, n_tiles, rows, 256, it, mode);
        cudaEventRecord(a); for (int it = 0; it < 10; it++) k_lines<<<148 * ctas, 128>>>(d, n_tiles, rows, 256, it, mode); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b); printf("whole-line spans, edge line by %s %.3f ms  %.1f GB/s\n", mode ? "8 lanes" : "1 lane x 8", ms / 10, bytes / (ms / 10) / 1e6);
    }
    return 0;
}
