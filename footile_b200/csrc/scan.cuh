// scan.cuh — generic 3-phase exclusive scan
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / EdgeRec / ...).
#pragma once

// ---------------------------------------------------------------------------
// generic 3-phase scan (reduce / scan partials / apply), exclusive, n+1 outputs
// ---------------------------------------------------------------------------
struct AddU32 {
    typedef uint32_t T;
    static __device__ __forceinline__ T identity() { return 0u; }
    static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
    static __device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
};
struct SumHeadOp {
    typedef SumHead T;
    static __device__ __forceinline__ T identity() { return {0u, NONE32}; }
    // the vertex total saturates instead of wrapping: more than 2^32 - 1 vertices then read as "does not fit" (cap_v
    // is at most 0x7FFFFFFF), never as a small wrapped count that would pass the capacity guard
    static __device__ __forceinline__ T combine(T a, T b) {
        uint32_t s = a.sum + b.sum;
        if (s < a.sum) s = 0xFFFFFFFFu;
        return {s, b.head != NONE32 ? a.sum + b.head : a.head};
    }
    static __device__ __forceinline__ T shfl_up(T v, int d) {
        return {__shfl_up_sync(0xFFFFFFFFu, v.sum, d), __shfl_up_sync(0xFFFFFFFFu, v.head, d)};
    }
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

// Exclusive scan of one value per thread across the block; returns the
// exclusive prefix and the block total (to all threads).
template <class Op>
__device__ typename Op::T block_exclusive(typename Op::T v, typename Op::T *total) {
    typedef typename Op::T T;
    __shared__ T warp_tot[SCAN_THREADS / 32];
    __shared__ T blk_tot;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = Op::shfl_up(inc, d);
        if (lane >= d) inc = Op::combine(o, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < SCAN_THREADS / 32 ? warp_tot[lane] : Op::identity();
        T winc = w;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            T o = Op::shfl_up(winc, d);
            if (lane >= d) winc = Op::combine(o, winc);
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = winc;  // inclusive over warps
        if (lane == SCAN_THREADS / 32 - 1) blk_tot = winc;
    }
    __syncthreads();
    T excl_in_warp = Op::shfl_up(inc, 1);
    if (lane == 0) excl_in_warp = Op::identity();
    T base = wid > 0 ? warp_tot[wid - 1] : Op::identity();
    *total = blk_tot;
    T r = Op::combine(base, excl_in_warp);
    __syncthreads();
    return r;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce(const typename Op::T *in, uint32_t n, typename Op::T *partials) {
    typedef typename Op::T T;
    uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    T acc = Op::identity();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) acc = Op::combine(acc, in[base + i]);
    T tot;
    block_exclusive<Op>(acc, &tot);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_partials(typename Op::T *partials, uint32_t n_blocks) {
    typedef typename Op::T T;
    T carry = Op::identity();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += SCAN_THREADS) {
        uint32_t i = b0 + threadIdx.x;
        T v = i < n_blocks ? partials[i] : Op::identity();
        T tot;
        T ex = block_exclusive<Op>(v, &tot);
        if (i < n_blocks) partials[i] = Op::combine(carry, ex);
        carry = Op::combine(carry, tot);
    }
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const typename Op::T *in, uint32_t n, const typename Op::T *partials,
                                                           typename Op::T *out) {
    typedef typename Op::T T;
    uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = Op::identity();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < n ? in[base + i] : Op::identity();
        acc = Op::combine(acc, v[i]);
    }
    T tot;
    T ex = block_exclusive<Op>(acc, &tot);
    T run = Op::combine(partials[blockIdx.x], ex);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run = Op::combine(run, v[i]);
        if (base + i + 1 == n) out[n] = run;
    }
}

// Small inputs: one block does the whole scan (a chunk of SCAN_BLOCK items per iteration, the running total
// carried in a register) instead of three launches, and hands the grand total to `fin` (thread 0) — the
// capacity guards set_vertex_count / set_entry_count ride along, so a replay saves three or four kernel
// boundaries per scan.  Same outputs as the 3-phase scan: exclusive, n + 1 values.
template <class Op, class Fin>
__global__ void __launch_bounds__(SCAN_THREADS) scan_small(const typename Op::T *in, uint32_t n, typename Op::T *out, Fin fin) {
    typedef typename Op::T T;
    T carry = Op::identity();
    for (uint32_t base0 = 0; base0 < n; base0 += SCAN_BLOCK) {
        const uint32_t base = base0 + threadIdx.x * SCAN_ITEMS;
        T v[SCAN_ITEMS];
        T acc = Op::identity();
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            v[i] = base + i < n ? in[base + i] : Op::identity();
            acc = Op::combine(acc, v[i]);
        }
        T tot;
        T ex = block_exclusive<Op>(acc, &tot);
        T run = Op::combine(carry, ex);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            if (base + i < n) out[base + i] = run;
            run = Op::combine(run, v[i]);
        }
        carry = Op::combine(carry, tot);
    }
    if (threadIdx.x == 0) {
        out[n] = carry;
        fin(carry);
    }
}
constexpr uint32_t SCAN_SMALL_MAX = 4 * SCAN_BLOCK;

