// stroke_kernels.cuh — the stroker on the device (stroker.rs:204-416): outline ops of flattened wide polylines
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / SumHead / ...).
//
// The reference walks each sub-stroke sequentially, once forward and once in reverse, and emits per segment an offset
// point or a join.  Segment i of a side needs only its own two points and the point before them, so here every
// (kept point, side) pair is a thread: a counting pass, a scan over fixed slots in the reference's output order, and
// an emitting pass that writes Line / Close ops straight into the op buffer of the fill pipeline — the outline never
// leaves HBM.  Arithmetic is the host stroker's (stroker.cpp), f32 op for op, with libm restated in libm_compat.cuh;
// a join whose |sin| test is too close to its threshold to predict glibc's sinf raises `fallback`, and the call is
// redone by the host stroker (nothing was drawn).
#pragma once

struct StrokeSub {  // host-built from the ops: a sub-stroke is a run of drawing ops between Move / Close (stroker.rs:204-236)
    uint32_t op_first, op_end;  // drawing ops [op_first, op_end) (PenWidth ops in between carry no points)
    uint32_t joined;            // the last close() applied to it was close(true) (stroker.rs:230-236)
    uint32_t job;
};
struct StrokeSubInfo {  // device-built: where the sub-stroke's kept points are
    uint32_t kstart, n, joined, job;
};
struct StrokeCounters {
    uint32_t n_raw;     // flattened points before Stroke::add_point's de-dup
    uint32_t nk;        // points kept
    uint32_t n_out;     // outline ops
    uint32_t overflow;  // n_raw exceeded the speculative capacity: nothing downstream ran
    uint32_t fallback;  // the host stroker must take this call (undecidable sin comparison, or a job at the 65 535-point cap)
    uint32_t need_raw;
    uint32_t pad[2];
};
struct StrokeJoin {
    int join;
    float miter_limit, tol_sq;
};

__global__ void stroke_set_raw(StrokeCounters *C, const SumHead *__restrict__ off, const JobDesc *__restrict__ jobs, uint32_t n_jobs, uint32_t n_ops,
                               uint32_t cap_raw) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) {
        const uint32_t n = off[n_ops].sum;
        C->n_raw = n;
        if (n > cap_raw) {
            C->overflow = 1;
            C->need_raw = n;
        }
    }
    // Stroke::add_point ignores points once 65 535 are stored (stroker.rs:206): such a stroke goes to the host stroker
    if (j < n_jobs && off[jobs[j].op_end].sum - off[jobs[j].op_begin].sum >= 65535u) C->fallback = 1;
}

// Stroke::add_point (stroker.rs:204-225): the first point after a close is always stored, any other point only when it
// differs (f32 ==) from the point before it.  keep[] covers the whole capacity so that the scan needs no device-side count.
__global__ void __launch_bounds__(256) stroke_keep(const float *__restrict__ xyw, const uint32_t *__restrict__ wop, const SumHead *__restrict__ off,
                                                   const uint32_t *__restrict__ opsub, const StrokeSub *__restrict__ subs,
                                                   const StrokeCounters *__restrict__ C, uint32_t cap_raw, uint32_t *__restrict__ keep) {
    const uint32_t n_raw = C->overflow ? 0u : C->n_raw;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < cap_raw; r += gridDim.x * blockDim.x) {
        uint32_t k = 0;
        if (r < n_raw) {
            const uint32_t op = wop[r];
            const bool first = r == off[op].sum && subs[opsub[op]].op_first == op;
            k = first || r == 0 || xyw[3 * (size_t)r] != xyw[3 * (size_t)r - 3] || xyw[3 * (size_t)r + 1] != xyw[3 * (size_t)r - 2];
        }
        keep[r] = k;
    }
}

__global__ void __launch_bounds__(256) stroke_compact(const float *__restrict__ xyw, const uint32_t *__restrict__ wop, const uint32_t *__restrict__ opsub,
                                                      const uint32_t *__restrict__ keep, const uint32_t *__restrict__ kidx,
                                                      StrokeCounters *__restrict__ C, uint32_t cap_raw, float4 *__restrict__ kp) {
    const uint32_t n_raw = C->overflow ? 0u : C->n_raw;
    if (blockIdx.x == 0 && threadIdx.x == 0) C->nk = kidx[cap_raw];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_raw; r += gridDim.x * blockDim.x)
        if (keep[r]) kp[kidx[r]] = make_float4(xyw[3 * (size_t)r], xyw[3 * (size_t)r + 1], xyw[3 * (size_t)r + 2], __uint_as_float(opsub[wop[r]]));
}

__global__ void stroke_sub_info(const StrokeSub *__restrict__ subs, uint32_t n_subs, const SumHead *__restrict__ off, const uint32_t *__restrict__ kidx,
                                const StrokeCounters *__restrict__ C, StrokeSubInfo *__restrict__ info) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_subs || C->overflow) return;
    const StrokeSub sb = subs[s];
    const uint32_t k0 = kidx[off[sb.op_first].sum], k1 = kidx[off[sb.op_end].sum];
    info[s] = {k0, k1 - k0, sb.joined, sb.job};
}

// ---- the outline of one segment (stroker.rs:265-416) ----
template <bool EMIT>
struct StrokeSink {
    uint32_t n = 0;
    ftl_path_op *out = nullptr;
    __device__ __forceinline__ void line(pointy::Pt p) {  // stroke_point (stroker.rs:312-314)
        if (EMIT) {
            float *q = reinterpret_cast<float *>(out + n);
            reinterpret_cast<uint32_t *>(q)[0] = FTL_OP_LINE;
            q[1] = p.x; q[2] = p.y; q[3] = 0.0f; q[4] = 0.0f; q[5] = 0.0f; q[6] = 0.0f;
        }
        n++;
    }
    __device__ __forceinline__ void close_op() {
        if (EMIT) {
            float *q = reinterpret_cast<float *>(out + n);
            reinterpret_cast<uint32_t *>(q)[0] = FTL_OP_CLOSE;
            q[1] = 0.0f; q[2] = 0.0f; q[3] = 0.0f; q[4] = 0.0f; q[5] = 0.0f; q[6] = 0.0f;
        }
        n++;
    }
};

__device__ __forceinline__ pointy::Pt stroke_normalize(pointy::Pt v) {  // pointy Pt::normalize: v / hypot, zero stays zero
    const float m = libm::hypotf_glibc(v.x, v.y);
    if (m > 0.0f) return {v.x / m, v.y / m};
    return {0.0f, 0.0f};
}
__device__ __forceinline__ float stroke_angle_rel(pointy::Pt a, pointy::Pt b) {  // pointy Pt::angle_rel, wrapped to (-pi, pi]
    const float pi = 3.14159265358979323846f;
    const float th = libm::atan2f_glibc(a.y, a.x) - libm::atan2f_glibc(b.y, b.x);
    if (th < -pi) return th + 2.0f * pi;
    if (th > pi) return th - 2.0f * pi;
    return th;
}
__device__ __forceinline__ bool stroke_intersection(pointy::Pt a0, pointy::Pt a1, pointy::Pt b0, pointy::Pt b1, pointy::Pt *out) {  // pointy Line::intersection
    const pointy::Pt av = pointy::sub(a0, a1), bv = pointy::sub(b0, b1);
    const float den = pointy::cross(av, bv);
    if (den == 0.0f) return false;
    const float ca = pointy::cross(a0, a1), cb = pointy::cross(b0, b1);
    const float xn = bv.x * ca - av.x * cb;
    const float yn = bv.y * ca - av.y * cb;
    *out = {xn / den, yn / den};
    return true;
}
// stroke_offset (stroker.rs:301-309)
__device__ __forceinline__ void stroke_offset(const float4 &p0, const float4 &p1, pointy::Pt *r0, pointy::Pt *r1) {
    const pointy::Pt pp0 = {p0.x, p0.y}, pp1 = {p1.x, p1.y};
    const pointy::Pt vr = stroke_normalize(pointy::right(pointy::sub(pp1, pp0)));
    *r0 = pointy::add(pp0, pointy::scale(vr, p0.z / 2.0f));
    *r1 = pointy::add(pp1, pointy::scale(vr, p1.z / 2.0f));
}
// stroke_arc (stroker.rs:399-416), depth-first with an explicit stack of the pending right halves
template <bool EMIT>
__device__ void stroke_arc(const float4 &p, pointy::Pt a, pointy::Pt b, float tol_sq, StrokeSink<EMIT> &sink) {
    pointy::Pt sc[MAX_DEPTH], sb[MAX_DEPTH];
    int sd[MAX_DEPTH];
    int sp = 0, depth = 0;
    const pointy::Pt pc = {p.x, p.y};
    for (;;) {
        const pointy::Pt vr = stroke_normalize(pointy::right(pointy::sub(b, a)));
        const pointy::Pt c = pointy::add(pc, pointy::scale(vr, p.z / 2.0f));
        const pointy::Pt ab = pointy::midpoint(a, b);
        if (pointy::distance_sq(c, ab) <= tol_sq || depth >= MAX_DEPTH) {
            sink.line(b);
            if (sp == 0) break;
            sp--;
            a = sc[sp]; b = sb[sp]; depth = sd[sp];
        } else {
            sc[sp] = c; sb[sp] = b; sd[sp] = depth + 1;
            sp++;
            b = c;
            depth++;
        }
    }
}
// stroke_join (stroker.rs:323-396); returns false when a sin comparison could not be decided
template <bool EMIT>
__device__ bool stroke_join(const StrokeJoin &sj, const float4 &p, pointy::Pt a0, pointy::Pt a1, pointy::Pt b0, pointy::Pt b1, StrokeSink<EMIT> &sink) {
    if (sj.join == FTL_JOIN_MITER) {
        const float ml = sj.miter_limit;
        if (ml > 0.0f) {
            const float sm_min = 1.0f / ml;
            const float th = stroke_angle_rel(pointy::sub(a1, a0), pointy::sub(b0, b1));
            const int ge = libm::abs_sin_ge(th / 2.0f, sm_min), lt = libm::abs_sin_lt_one(th / 2.0f);
            if (ge < 0 || (ge == 1 && lt < 0)) return false;
            pointy::Pt xp;
            if (ge == 1 && lt == 1 && stroke_intersection(a0, a1, b0, b1, &xp)) {
                sink.line(xp);
                return true;
            }
        }
        sink.line(a1);
        sink.line(b0);
    } else if (sj.join == FTL_JOIN_BEVEL) {
        sink.line(a1);
        sink.line(b0);
    } else {
        const float th = stroke_angle_rel(pointy::sub(a1, a0), pointy::sub(b0, b1));
        sink.line(a1);
        if (th <= 0.0f) sink.line(b0);
        else stroke_arc<EMIT>(p, a1, b0, sj.tol_sq, sink);
    }
    return true;
}

// Output order of the reference (stroker.rs:250-262): per sub-stroke the forward side, Close if joined, the reverse
// side, Close.  Slot of segment i of a side: 2 * (kstart + s) + side * (n + 1) + i  (i <= n).
__device__ __forceinline__ uint32_t stroke_slot(const StrokeSubInfo &si, uint32_t s, uint32_t side, uint32_t i) { return 2u * (si.kstart + s) + side * (si.n + 1u) + i; }

template <bool EMIT>
__device__ void stroke_segment(const StrokeJoin &sj, const float4 *__restrict__ kp, const StrokeSubInfo &si, uint32_t s, uint32_t side, uint32_t i,
                               uint32_t len, uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off, ftl_path_op *__restrict__ ops_out,
                               StrokeCounters *C) {
    const uint32_t n = si.n, k0 = si.kstart;
    uint32_t v0, v1, vp;
    if (side == 0) {  // SubStroke::next Forward from `start` (stroker.rs:67-76)
        v0 = i < n ? i : 0u;
        v1 = v0 + 1 < n ? v0 + 1 : 0u;
        vp = i - 1;  // only read when i >= 1
    } else {  // Reverse from sub_end (stroker.rs:77-84,145-149)
        v0 = i < n ? n - 1 - i : n - 1;
        v1 = v0 > 0 ? v0 - 1 : n - 1;
        vp = n - i;  // i >= 1
    }
    const float4 p0 = kp[k0 + v0], p1 = kp[k0 + v1];
    pointy::Pt r0, r1;
    stroke_offset(p0, p1, &r0, &r1);
    StrokeSink<EMIT> sink;
    const uint32_t slot = stroke_slot(si, s, side, i);
    if (EMIT) sink.out = ops_out + off[slot];
    if (i >= 1) {
        pointy::Pt x0, x1;
        stroke_offset(kp[k0 + vp], p0, &x0, &x1);
        if (!stroke_join<EMIT>(sj, p0, x0, x1, r0, r1, sink) && !EMIT) atomicOr(&C->fallback, 1u);
    } else if (!si.joined) sink.line(r0);
    if (i + 1 == len) {
        if (!si.joined) sink.line(r1);
        if (side == 1 || si.joined) sink.close_op();
    }
    if (!EMIT) cnt[slot] = sink.n;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) stroke_segments(StrokeJoin sj, const float4 *__restrict__ kp, const StrokeSubInfo *__restrict__ info,
                                                       StrokeCounters *C, uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off,
                                                       ftl_path_op *__restrict__ ops_out) {
    if (C->overflow || (EMIT && C->fallback)) return;
    const uint32_t nk = C->nk;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * nk; t += gridDim.x * blockDim.x) {
        const uint32_t k = t >> 1, side = t & 1u;
        const uint32_t s = __float_as_uint(kp[k].w);
        const StrokeSubInfo si = info[s];
        const uint32_t j = k - si.kstart, n = si.n;
        const uint32_t len = si.joined ? n + 1 : n - 1;  // SubStroke::len (stroker.rs:88-96); n >= 1
        if (len == 0) continue;
        const uint32_t i = side == 0 ? j : n - 1 - j;
        if (i < len) stroke_segment<EMIT>(sj, kp, si, s, side, i, len, cnt, off, ops_out, C);
        if (si.joined && (side == 0 ? j == 0 : j == n - 1)) stroke_segment<EMIT>(sj, kp, si, s, side, n, len, cnt, off, ops_out, C);  // the closing segment
    }
}

__global__ void stroke_set_out(StrokeCounters *C, const uint32_t *__restrict__ off, uint32_t n_slots) { C->n_out = C->overflow ? 0u : off[n_slots]; }

// Op ranges of the outline per job, written into the job descriptors of the fill that follows.
__global__ void stroke_patch_jobs(JobDesc *__restrict__ jobs, uint32_t n_jobs, const uint32_t *__restrict__ job_first_sub, uint32_t n_subs,
                                  const StrokeSubInfo *__restrict__ info, const uint32_t *__restrict__ off, const StrokeCounters *__restrict__ C) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const uint32_t s0 = job_first_sub[j], s1 = job_first_sub[j + 1];
    const uint32_t total = C->n_out;
    jobs[j].op_begin = s0 < n_subs ? off[2u * (info[s0].kstart + s0)] : total;
    jobs[j].op_end = s1 < n_subs ? off[2u * (info[s1].kstart + s1)] : total;
}
