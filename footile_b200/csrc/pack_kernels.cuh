// pack_kernels.cuh — packed read-back and raster checksums
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / EdgeRec / ...).
#pragma once

// ---------------------------------------------------------------------------
// packed read-back: rasters are mostly long constant spans, and PCIe is ~100x
// slower than HBM, so device->host copies of large rasters travel as
//   code[b]   : the byte value of 32-byte block b if the block is uniform
//   bitmap[u] : bit i set = block 32u+i is literal (not uniform)
//   off[u]    : literal blocks before unit u (exclusive scan of the popcounts)
//   literals  : the literal blocks, 32 bytes each, in order
// and are expanded into the caller's buffer by host threads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_classify(const uint4 *__restrict__ src, size_t n_blocks, uint8_t *__restrict__ code,
                                                     uint32_t *__restrict__ bitmap, uint32_t *__restrict__ cnt) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // n_blocks is a multiple of 32: warps are full or empty
    if (b >= n_blocks) return;
    const uint4 lo = src[2 * b], hi = src[2 * b + 1];
    const uint32_t v = (lo.x & 0xFFu) * 0x01010101u;
    const bool uniform = lo.x == v && lo.y == v && lo.z == v && lo.w == v && hi.x == v && hi.y == v && hi.z == v && hi.w == v;
    code[b] = (uint8_t)(v & 0xFFu);
    const uint32_t lit = __ballot_sync(0xFFFFFFFFu, !uniform);
    if ((threadIdx.x & 31) == 0) {
        bitmap[b >> 5] = lit;
        cnt[b >> 5] = __popc(lit);
    }
}
__global__ void __launch_bounds__(256) pack_literals(const uint4 *__restrict__ src, size_t n_blocks, const uint32_t *__restrict__ bitmap,
                                                     const uint32_t *__restrict__ off, uint4 *__restrict__ lit) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t m = bitmap[b >> 5], lane = threadIdx.x & 31;
    if ((m >> lane) & 1u) {
        const size_t k = (size_t)off[b >> 5] + __popc(m & ((1u << lane) - 1u));
        lit[2 * k] = src[2 * b];
        lit[2 * k + 1] = src[2 * b + 1];
    }
}

// 64-bit FNV-1a per raster (parity checks of large batches): one CTA per
// raster hashes 256 interleaved lanes, then lane digests are folded in order.
__global__ void __launch_bounds__(256) fnv_rasters(const uint8_t *__restrict__ base, size_t raster_bytes, uint64_t *__restrict__ out) {
    __shared__ uint64_t part[256];
    const uint8_t *p = base + (size_t)blockIdx.x * raster_bytes;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = threadIdx.x; i < raster_bytes; i += 256) h = (h ^ p[i]) * 0x100000001b3ull;
    part[threadIdx.x] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t g = 0xcbf29ce484222325ull;
        for (int i = 0; i < 256; i++)
            for (int b = 0; b < 8; b++) g = (g ^ ((part[i] >> (8 * b)) & 0xFF)) * 0x100000001b3ull;
        out[blockIdx.x] = g;
    }
}

// ---------------------------------------------------------------------------
// Output conversion (examples/fishy.rs:33, examples/png/mod.rs:22-27): the raster the reference hands to its PNG
// encoder is Raster::<SRgba8>::with_raster(&p.raster()) - every Rgba8p pixel (linear, premultiplied) becomes SRgba8
// (sRGB gamma, straight alpha): colour / alpha in Ch8, then the sRGB transfer function; alpha is copied.  Graya8p ->
// SGraya8 likewise; a Matte8 raster is reinterpreted as SGray8 byte for byte (png/mod.rs:22-27: no arithmetic).
// pix_compat.cuh isolates the RECALLED pix arithmetic (Ch8 division = min((c << 8) / a, 255), 0 for a = 0; the encode
// table is round(255 * srgb(i / 255))).  One 16-byte word per thread: a pure streaming pass.
// ---------------------------------------------------------------------------
__constant__ uint8_t c_srgb_encode[256];
template <int FMT>
__global__ void __launch_bounds__(256) srgb_convert(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n_words) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint4 v = src[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (FMT == FTL_RGBA8P) {
            const uint32_t a = w[k] >> 24;
            uint32_t o = a << 24;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) o |= (uint32_t)c_srgb_encode[pix::ch8_div((w[k] >> (8 * ch)) & 0xFFu, a)] << (8 * ch);
            w[k] = o;
        } else {  // two (gray, alpha) pixels per word
            uint32_t o = 0;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const uint32_t px = (w[k] >> (16 * e)) & 0xFFFFu, a = px >> 8;
                o |= ((uint32_t)c_srgb_encode[pix::ch8_div(px & 0xFFu, a)] | (a << 8)) << (16 * e);
            }
            w[k] = o;
        }
    }
    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
