// ctx.cpp — several GPUs behind one C handle: independent paths sharded over the devices, or one raster split
// into row bands, one host thread per device, no data-path collective.  Built only on the public C ABI above it, so a Rust
// or C caller gets exactly what footile_b200/sharding.py gives the Python mirror.
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "engine.h"

using namespace ftl;

struct ftl_ctx {
    std::vector<int> devices;
};

namespace {
int bad(const char *msg) {
    set_error(msg);
    return FTL_ERR_INVALID;
}
// Run fn(rank) on one thread per device; the first failure (lowest rank) is reported with its message.
template <class F>
int on_all(const ftl_ctx *c, F fn) {
    const size_t n = c->devices.size();
    std::vector<int> rc(n, FTL_OK);
    std::vector<std::string> msg(n);
    auto body = [&](size_t r) {
        rc[r] = fn((uint32_t)r);
        if (rc[r]) msg[r] = last_error();  // thread-local: copy it out of the worker
    };
    if (n == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (size_t r = 0; r < n; r++) th.emplace_back(body, r);
        for (std::thread &t : th) t.join();
    }
    for (size_t r = 0; r < n; r++)
        if (rc[r]) {
            set_error("device " + std::to_string(c->devices[r]) + ": " + msg[r]);
            return rc[r];
        }
    return FTL_OK;
}
}  // namespace

extern "C" {

int ftl_shard_range(uint32_t n, uint32_t rank, uint32_t world, uint32_t *first, uint32_t *count) {
    if (!first || !count || world == 0 || rank >= world) return bad("bad shard arguments");
    const uint32_t base = n / world, extra = n % world;
    *first = rank * base + std::min(rank, extra);
    *count = base + (rank < extra ? 1u : 0u);
    return FTL_OK;
}

int ftl_band_rows(uint32_t height, uint32_t rank, uint32_t world, uint32_t align, uint32_t *row_begin, uint32_t *row_end) {
    if (!row_begin || !row_end || world == 0 || rank >= world) return bad("bad band arguments");
    if (align == 0) align = 32;
    const uint32_t units = (height + align - 1) / align;
    uint32_t first, count;
    ftl_shard_range(units, rank, world, &first, &count);
    *row_begin = std::min(first * align, height);
    *row_end = std::min((first + count) * align, height);
    return FTL_OK;
}

int ftl_ctx_new(int n_devices, const int *devices, ftl_ctx **out) {
    try {
        if (!out) return bad("out is null");
        *out = nullptr;
        int have = 0;
        int rc = ftl_device_count(&have);
        if (rc) return rc;
        if (n_devices <= 0) n_devices = have;  // all of them
        ftl_ctx *c = new ftl_ctx();
        for (int i = 0; i < n_devices; i++) {
            const int d = devices ? devices[i] : i;
            if (d < 0 || d >= have) {
                delete c;
                return bad("device index out of range");
            }
            c->devices.push_back(d);  // a device may appear more than once (several shards on one GPU)
        }
        *out = c;
        return FTL_OK;
    } catch (...) {
        set_error("host allocation failed");
        return FTL_ERR_NOMEM;
    }
}
int ftl_ctx_free(ftl_ctx *c) {
    delete c;
    return FTL_OK;
}
int ftl_ctx_size(const ftl_ctx *c) { return c ? (int)c->devices.size() : 0; }

int ftl_ctx_fill_batch(ftl_ctx *c, uint32_t width, uint32_t height, int format, float tolerance, uint32_t n_jobs, const ftl_path_op *ops,
                       const uint64_t *op_offsets, const uint8_t *rules, const float *transforms, const uint8_t *colors, void *dst, size_t nbytes) {
    try {
        if (!c || !op_offsets || (!dst && nbytes)) return bad("null argument");
        const size_t bpp = format == FTL_MATTE8 ? 1 : (format == FTL_GRAYA8P ? 2 : 4);
        const size_t raster = (size_t)width * height * bpp;
        if (nbytes != raster * n_jobs) return bad("nbytes does not match n_jobs*height*width*bpp");
        const uint32_t world = (uint32_t)c->devices.size();
        return on_all(c, [&](uint32_t r) -> int {
            uint32_t first, count;
            ftl_shard_range(n_jobs, r, world, &first, &count);
            if (count == 0) return FTL_OK;
            ftl_batch *b = nullptr;
            int rc = ftl_batch_new(width, height, format, count, c->devices[r], &b);
            if (rc) return rc;
            if (tolerance > 0.0f) ftl_batch_set_tolerance(b, tolerance);
            std::vector<uint64_t> offs(count + 1);  // the shard's offsets, rebased to its own first op
            for (uint32_t j = 0; j <= count; j++) offs[j] = op_offsets[first + j] - op_offsets[first];
            rc = ftl_batch_fill(b, count, ops + op_offsets[first], offs.data(), rules ? rules + first : nullptr,
                                transforms ? transforms + 6 * (size_t)first : nullptr, colors ? colors + 4 * (size_t)first : nullptr);
            if (!rc) rc = ftl_batch_read(b, 0, count, (uint8_t *)dst + raster * first, raster * count);
            std::string keep = rc ? last_error() : "";
            ftl_batch_free(b);
            if (rc) set_error(keep);
            return rc;
        });
    } catch (...) {
        set_error("host allocation failed");
        return FTL_ERR_NOMEM;
    }
}

int ftl_ctx_fill_bands(ftl_ctx *c, uint32_t width, uint32_t height, int format, int rule, const ftl_path_op *ops, size_t n_ops, const float transform[6],
                       float tolerance, const uint8_t *color, const void *init_pixels, void *dst, size_t nbytes) {
    try {
        if (!c || (!dst && nbytes)) return bad("null argument");
        const size_t bpp = format == FTL_MATTE8 ? 1 : (format == FTL_GRAYA8P ? 2 : 4);
        const size_t pitch = (size_t)width * bpp;
        if (nbytes != pitch * height) return bad("nbytes does not match height*width*bpp");
        const uint32_t world = (uint32_t)c->devices.size();
        return on_all(c, [&](uint32_t r) -> int {
            uint32_t r0, r1;
            ftl_band_rows(height, r, world, 32, &r0, &r1);
            if (r0 >= r1) return FTL_OK;
            ftl_plotter *p = nullptr;
            int rc = ftl_plotter_new_band(width, height, r0, r1, format, init_pixels ? (const uint8_t *)init_pixels + pitch * r0 : nullptr, c->devices[r], &p);
            if (rc) return rc;
            if (tolerance > 0.0f) ftl_set_tolerance(p, tolerance);
            if (transform) rc = ftl_set_transform(p, transform);
            if (!rc) rc = ftl_fill(p, rule, ops, n_ops, color);
            if (!rc) rc = ftl_read_raster(p, (uint8_t *)dst + pitch * r0, pitch * (r1 - r0));
            std::string keep = rc ? last_error() : "";
            ftl_plotter_free(p);
            if (rc) set_error(keep);
            return rc;
        });
    } catch (...) {
        set_error("host allocation failed");
        return FTL_ERR_NOMEM;
    }
}

}  // extern "C"
