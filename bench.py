#!/usr/bin/env python
"""bench.py — throughput of the footile hot path on B200 (and the CPU baseline beside it).

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ...]
A step is one pass of the hot path over one batch of synthetic paths.  Default workload =
BASELINE.json configs[1]: heptagram fills (NonZero / EvenOdd alternating) into 4096x4096 Matte8
rasters, `--batch` rasters per step (default 256 = 4 GiB of output, far beyond the 126 MB L2).
  value : Gpx/s with the path ops already resident in HBM (ftl_batch_run / ftl_fill_replay), CUDA events
          on the library's stream, max over ranks.
  e2e   : the same metric through the C ABI with HOST buffers: ftl_batch_fill(host ops) (strokes4k:
          ftl_batch_stroke, the stroker inside the timed region) then ftl_batch_read of every raster into
          pinned host memory.  d2h_bytes_per_step is the logical size of the rasters returned,
          wire_*_bytes_per_step what really crossed PCIe (the read-back travels packed).
  roofline : the tile kernels (raster_tiles + raster_bins), algorithmic bytes / CUDA-event duration vs
          MEASURED_PEAKS.json; `traffic` is the ncu DRAM byte count of one launch of this workload,
          taken from profiles/traffic.json (`traffic_source` says so: it is not measured in this run).
  cpu_baseline : the oracle (CPU restatement of footile, 1 thread) on a bounded sample, rank 0, N=1.
  secondary (default workload only): the two multi-GPU configs of BASELINE.json at this N -
          config 5 (one 32768^2 raster split into row bands, strong scaling) and config 4 (a 4096-path
          slice of the 100k-path batch per GPU, weak scaling) - and the per-call latency of
          benches/fishyb.rs's fills.
--impl reference times the oracle on all host threads (one plotter per thread).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PIXELS_NOTE = "px = W * (H - max(top_row,0)) per fill: the reference resolves every pixel of those rows (fig.rs:497,539)"
DTYPE = "i32 fixed-point 16.16 / u8"
WORKLOADS = ["heptagram", "batch512", "fishy256", "strokes4k", "bigraster", "latency"]
LABEL = {"heptagram": "heptagram fill 4096^2", "batch512": "100k-path batch 512^2", "fishy256": "fishyb fill_256 batch 256^2",
         "strokes4k": "stroke scenes x30 3840x2160", "bigraster": "one 32768^2 raster"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def recorded_traffic(key):
    """DRAM bytes per launch of the tile kernel of this workload from the committed ncu capture (None if there is none)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            v = json.load(f).get(key)
        return (float(v), "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch)") if v else (None, None)
    except Exception:
        return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a thread every 2 ms (the
    timed region of the default run is ~20 ms, too short for `nvidia-smi -lms`)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.h, self.sm, self.mask, self.mx, self.run, self.t = index, None, [], 0, None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.run = True
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception:
            self.h = None
        return self

    def _loop(self):
        while self.run:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.run = False
        self.t.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(n for n, bit in self.REASONS.items() if self.mask & bit), "samples": len(self.sm)}


class HostBuffer:
    """Page-locked host buffer for the step's results, backed by transparent huge pages where the kernel grants
    them (mmap + MADV_HUGEPAGE + cudaHostRegister): the packed read-back is bound by the host threads'
    streaming stores, and 2 MiB pages spare them a page walk every 4 KiB."""

    def __init__(self, nbytes):
        import mmap
        self.n = int(nbytes)
        self.m = mmap.mmap(-1, self.n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        if os.environ.get("FTL_BENCH_NO_THP") != "1":
            try:
                self.m.madvise(mmap.MADV_HUGEPAGE)
            except Exception:
                pass
        self.a = np.frombuffer(self.m, dtype=np.uint8)
        self.a[:] = 0  # fault the pages in
        self.ptr = self.a.ctypes.data
        self.registered = False
        try:
            import torch
            self.registered = int(torch.cuda.cudart().cudaHostRegister(self.ptr, self.n, 0)) == 0
        except Exception:
            pass

    def close(self):
        """Unregister BEFORE the pages are unmapped: a stale registration makes later allocations fail ("already mapped")."""
        if self.registered:
            try:
                import torch
                torch.cuda.cudart().cudaHostUnregister(self.ptr)
            except Exception:
                pass
            self.registered = False

    def __del__(self):
        self.close()

    def data_ptr(self):
        return self.ptr

    def numel(self):
        return self.n


# ---- workloads -----------------------------------------------------------------
def make_workload(name, batch, rank):
    """Synthetic input of one step: dict(ops, offs, rules, tr, size | w,h, desc [, stroke=source paths of config 3])."""
    from footile_b200 import scenes
    if name == "heptagram":
        size = 4096
        path = scenes.heptagram_abs()
        return dict(size=size, ops=np.tile(path, batch), offs=np.arange(batch + 1, dtype=np.uint64) * np.uint64(len(path)),
                    rules=(np.arange(batch) & 1).astype(np.uint8), tr=np.tile(scenes.heptagram_transform(size), (batch, 1)),
                    desc="heptagram {7/2} fill, NonZero/EvenOdd alternating, 4096x4096 Matte8")
    if name == "batch512":
        ops, offs, rules = scenes.random_curve_paths(rank * batch, batch)
        return dict(size=512, ops=ops, offs=offs, rules=rules, tr=None, desc="random quad/cubic paths (64 segments), 512x512 Matte8 per path")
    if name == "fishy256":
        path = scenes.fishy_bench()
        return dict(size=256, ops=np.tile(path, batch), offs=np.arange(batch + 1, dtype=np.uint64) * np.uint64(len(path)),
                    rules=np.zeros(batch, dtype=np.uint8), tr=np.tile(np.array([2, 0, 0, 0, 2, 0], dtype=np.float32), (batch, 1)),
                    desc="benches/fishyb.rs fill_256: fishy path, scale(2,2), 256x256 Matte8 per fill")
    if name == "strokes4k":
        # config 3: the six stroke scenes x30 with Round joins, one stroke per 3840x2160 Rgba8p raster.  `stroke` holds the
        # SOURCE paths (what the user hands to Plotter::stroke); ops/offs are filled in by the caller with the outlines
        # (what Plotter::stroke hands to fill, plotter.rs:361-364) for the legs that replay resident geometry.
        src = list(scenes.stroke_scenes(30.0).values())
        parts = [src[j % len(src)] for j in range(batch)]
        soffs = np.zeros(batch + 1, dtype=np.uint64)
        soffs[1:] = np.cumsum([len(p) for p in parts])
        return dict(size=0, w=3840, h=2160, stroke=(np.ascontiguousarray(np.concatenate(parts)), soffs, src), rules=np.zeros(batch, dtype=np.uint8), tr=None,
                    desc="stroke.rs/stroke2.rs/round.rs/over.rs/teeth.rs/curve.rs x30, Round joins, butt ends: stroke into 3840x2160 Rgba8p")
    raise SystemExit("unknown workload " + name)


def set_outlines(wl, batch, outline_of):
    """config 3: outlines of the source paths through `outline_of` (the product's stroker in our arm, the oracle's in the reference arm)."""
    t0 = time.perf_counter()
    outs = [outline_of(p) for p in wl["stroke"][2]]
    wl["outline_ms"] = 1e3 * (time.perf_counter() - t0)
    parts = [outs[j % len(outs)] for j in range(batch)]
    wl["offs"] = np.zeros(batch + 1, dtype=np.uint64)
    wl["offs"][1:] = np.cumsum([len(p) for p in parts])
    wl["ops"] = np.ascontiguousarray(np.concatenate(parts))


def oracle_pixels(wl, sample):
    """Pixels per fill by the reference's dense-row convention, from the oracle's top_row (reference arm only)."""
    import oracle
    px = []
    W, H = wl.get("w", wl["size"]), wl.get("h", wl["size"])
    for j in sample:
        o = oracle.Plotter(8, 8, oracle.MATTE8)  # tiny raster: only (dir, top_row) are needed
        if wl["tr"] is not None:
            o.set_transform(wl["tr"][j])
        o.fill(int(wl["rules"][j]), wl["ops"][int(wl["offs"][j]): int(wl["offs"][j + 1])], (255,))
        px.append(W * max(0, H - max(o.last_info()["top_row"], 0)))
    return px


def cpu_run(wl, jobs, threads):
    """Seconds the oracle needs for `jobs` (indices) on `threads` C++ threads (rasters pre-allocated, timed in C++)."""
    import oracle
    W, H = wl.get("w", wl["size"]), wl.get("h", wl["size"])
    jobs = list(jobs)
    parts = [wl["ops"][int(wl["offs"][j]): int(wl["offs"][j + 1])] for j in jobs]
    offs = np.zeros(len(jobs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    ops = np.concatenate(parts)
    rules = np.array([wl["rules"][j] for j in jobs], dtype=np.uint8)
    tr = None if wl["tr"] is None else np.ascontiguousarray(np.array([wl["tr"][j] for j in jobs], dtype=np.float32))
    return oracle.batch_fill_timed(W, H, wl.get("ofmt", oracle.MATTE8), ops, offs, rules, tr, wl.get("color", (255,)), threads, 1)


class Dist:
    """The multi-process plumbing: one process per GPU (torchrun), NCCL only for barriers and the max over ranks."""

    def __init__(self, need_cuda=True):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.torch = None
        if need_cuda:
            import torch
            self.torch = torch
            if not torch.cuda.is_available():
                raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
            torch.cuda.set_device(self.local_rank)
            if self.world > 1:
                import torch.distributed as dist
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
                self.dist = dist

    def barrier(self, handle=None):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        if handle is not None:
            handle.sync()

    def max(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def done(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ---- config 5: one huge raster in row bands ------------------------------------------------
BIG = dict(size=32768, polys=160000)
BIG_METRIC = "Gpx/s composited (one 32768^2 Matte8 raster, 10.24 M edges, row bands)"


def big_config(world):
    return {"workload": "160 000 closed 64-gon sub-figures (seed 0xB160000^i) in ONE EvenOdd fill of a 32768x32768 Matte8 raster",
            "raster": "32768x32768", "format": "Matte8", "bands": world, "pixels": PIXELS_NOTE,
            "l2": "1 GiB raster and 328 MB of edges per fill: far beyond the 126 MB L2"}


def big_ops():
    from footile_b200 import scenes
    return scenes.random_polygons(0, BIG["polys"], vertices=64, size=BIG["size"], extent=2048)


def big_cpu_sample(ops, cores, repeats=1):
    """The order-free oracle on `cores` stripes of 16 rows, one per thread: (Gpx/s, description).  Every stripe walks all
    10.24 M edges (the reference's own loop would too: fig.rs:539 visits every row with its active-edge list)."""
    import oracle
    size, rows_each = BIG["size"], 16
    rng = np.random.default_rng(5)
    starts = [int(r) for r in rng.integers(0, size - rows_each, cores)]

    def stripe(r0):
        o = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True)
        o.set_rows(r0, r0 + rows_each)
        o.fill(1, ops, (255,))

    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        ths = [threading.Thread(target=stripe, args=(r0,)) for r0 in starts]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    return size * rows_each * cores / t / 1e9, t, "%d stripes of %d rows (order-free oracle, u32 vertex ids), one per thread on %d threads; every stripe walks all 10.24 M edges" % (cores, rows_each, cores)


def bigraster_ours(D, steps, warmup, ops=None, full=True):
    """One 32768x32768 Matte8 raster, 160 000 closed 64-gons (10.24 M edges) in one EvenOdd fill, rows split into
    `world` bands, one band per GPU, no collective: every rank holds the ops, culls the sub-figures outside its band on
    the device and recomputes (dir, top_row).  Strong scaling: the work is fixed, `value` = pixels of the whole raster
    / slowest rank's time.  Returns the fields of the JSON line (rank 0) or None."""
    import footile_b200 as fb
    from footile_b200 import Format, sharding
    torch = D.torch
    size = BIG["size"]
    ops = big_ops() if ops is None else ops
    r0, r1 = sharding.band_rows(size, D.rank, D.world, align=32)
    if os.environ.get("FTL_BENCH_BAND"):  # profiling aid: one GPU plays band R of W ("R/W"), e.g. for a launch list of an 8-GPU rank
        er, ew = (int(v) for v in os.environ["FTL_BENCH_BAND"].split("/"))
        r0, r1 = sharding.band_rows(size, er, ew, align=32)
    p = fb.Plotter.with_clear(size, size, Format.Matte8, device=D.local_rank, rows=(r0, r1))
    stream = torch.cuda.ExternalStream(p.stream(), device=D.local_rank)
    # ---- value: ops resident in HBM (ftl_fill_upload once, ftl_fill_replay per step) ----
    p.upload(1, ops, (255,))
    for _ in range(max(1, min(warmup, 3))):
        p.replay()
    D.barrier(p)
    sampler = ClockSampler(D.local_rank).start()
    fb.set_profiling(True)
    fb.tile_kernel_time(reset=True)
    l0 = fb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        p.replay()
    ev1.record(stream)
    D.barrier(p)
    ms = D.max(ev0.elapsed_time(ev1))
    launches = fb.launch_count() - l0
    tile_ms, tile_n = fb.tile_kernel_time(reset=True)
    fb.set_profiling(False)
    clocks = sampler.stop()
    top = p.debug_last_fill()["top_row"]
    px = size * (size - max(top, 0))
    out = {"ms_per_fill": ms / steps, "value": px * steps / (ms * 1e-3) / 1e9, "edges_per_s": BIG["polys"] * 64 * steps / (ms * 1e-3),
           "tile_ms_per_launch": tile_ms / max(tile_n, 1), "launches": int(launches), "clocks": clocks, "px": px, "rows": (r0, r1)}
    if full:
        # ---- e2e: host ops in (291 MB per fill), the band's rows back to pinned host memory ----
        pinned = HostBuffer((r1 - r0) * size)
        fb.transfer_bytes(reset=True)
        p.fill(1, ops, (255,))
        fb._lib.check(fb._lib.lib().ftl_read_raster(p._handle, pinned.data_ptr(), pinned.numel()))
        D.barrier(p)
        fb.transfer_bytes(reset=True)
        t0 = time.perf_counter()
        for _ in range(2):
            p.fill(1, ops, (255,))
            fb._lib.check(fb._lib.lib().ftl_read_raster(p._handle, pinned.data_ptr(), pinned.numel()))
        out["e2e_s"] = D.max((time.perf_counter() - t0) / 2)
        h2d, d2h = fb.transfer_bytes(reset=True)
        out["wire"] = (h2d // 2, d2h // 2)
        out["ops_bytes"] = int(ops.nbytes)
        pinned.close()
    del p
    return out


def run_bigraster(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    config = big_config(world)
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        v, t, sample = big_cpu_sample(big_ops(), cores, repeats=max(1, min(args.steps, 3)))
        print(json.dumps({"impl": "reference", "metric": BIG_METRIC, "value": v, "unit": "Gpx/s", "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": 0,
                          "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE,
                          "data": "synthetic", "config": config, "cpu_baseline": {"value": v, "unit": "Gpx/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    D = Dist()
    ops = big_ops()
    steps = max(1, args.steps)
    r = bigraster_ours(D, steps, args.warmup, ops=ops, full=not args.kernel_only)
    if D.rank == 0:
        peak, peak_kind = peaks()
        ach = r["px"] / D.world / (r["tile_ms_per_launch"] * 1e-3) / 1e9
        traffic, tsrc = recorded_traffic("bigraster")
        cpu = None
        if D.world == 1 and not args.kernel_only:
            v, t, sample = big_cpu_sample(ops, 1)
            cpu = {"value": v, "unit": "Gpx/s", "cores": 1, "kind": "port", "sample": sample.replace("one per thread on 1 threads", "one thread")}
        line = {"metric": BIG_METRIC, "value": r["value"], "unit": "Gpx/s", "n_gpus": D.world, "steps": steps, "warmup": args.warmup, "ms_per_step": r["ms_per_fill"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config,
                "edges_per_s": r["edges_per_s"], "clocks": r["clocks"], "gpu_launches": r["launches"],
                "timing": "CUDA events on the library stream around `steps` ftl_fill_replay calls (ops resident in HBM), max over ranks",
                "roofline": {"kernel": "raster_bins", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_kind": peak_kind,
                             "traffic": traffic, "traffic_source": tsrc, "avg_launch_ms": r["tile_ms_per_launch"],
                             "algorithmic_bytes_per_launch": r["px"] / D.world + 32.0 * BIG["polys"] * 64,
                             "note": "1 B per resolved pixel + 32 B per edge record; the kernel is instruction-bound, not bandwidth-bound: ~7 G (edge,row) "
                                     "items of ~90 warp-instructions per 32 per fill (profiles/README.md)"},
                "cpu_baseline": cpu}
        if "e2e_s" in r:
            line["e2e"] = {"value": r["px"] / r["e2e_s"] / 1e9, "unit": "Gpx/s", "h2d_bytes_per_step": r["ops_bytes"], "d2h_bytes_per_step": int((r["rows"][1] - r["rows"][0]) * BIG["size"]),
                           "wire_h2d_bytes_per_step": r["wire"][0], "wire_d2h_bytes_per_step": r["wire"][1], "ms_per_step": 1e3 * r["e2e_s"]}
        print(json.dumps(line))
    D.done()


# ---- per-call latency of small fills (benches/fishyb.rs) --------------------------------------------
def latency_numbers(device, cpu=True):
    import footile_b200 as fb
    from footile_b200 import Format, scenes
    out = {}
    path = scenes.fishy_bench()
    fish, eye = scenes.fishy_example()
    T = np.array([2, 0, 0, 0, 2, 0], dtype=np.float32)
    cases = [("fill_16", 16, Format.Matte8, T, path, (255,)), ("fill_256", 256, Format.Matte8, T, path, (255,)),
             ("fishy_fill_128_rgba8p", 128, Format.Rgba8p, None, fish, (127, 96, 96, 255))]
    for name, size, fmt, tr, ops, clr in cases:
        g = fb.Plotter(fb.Raster(size, size, fmt), device=device)
        if tr is not None:
            g.set_transform(tr)
        g.time_fills(0, ops, clr, iters=100)
        rec = {"gpu_us_per_call_sync_each": g.time_fills(0, ops, clr, iters=2000, sync_each=True),
               "gpu_us_per_call_back_to_back": g.time_fills(0, ops, clr, iters=2000, sync_each=False)}
        if cpu:
            import oracle
            ofmt = oracle.MATTE8 if fmt == Format.Matte8 else oracle.RGBA8P
            offs = np.array([0, len(ops)], dtype=np.uint64)
            reps = 2000
            t = oracle.batch_fill_timed(size, size, ofmt, ops, offs, np.zeros(1, dtype=np.uint8), None if tr is None else tr.reshape(1, 6), clr, 1, reps)
            rec["cpu_us_per_call"] = 1e6 * t / reps
        out[name] = rec
    out["what"] = ("ftl_fill on a resident plotter, timed inside the library (ftl_time_fills): with ftl_sync after every call, and 2000 calls back to back; "
                   "cpu = the oracle's fill of the same path into a pre-allocated raster, one thread, timed in C++")
    return out


# ---- batched workloads ------------------------------------------------------------------------------
def run_batch(D, args, name, batch, fmt, steps, warmup, full=True):
    """value / e2e / roofline / cpu_baseline of one batched workload; returns the JSON line as a dict (all ranks compute, rank 0 reports)."""
    import footile_b200 as fb
    from footile_b200 import Batch, Format
    torch = D.torch
    graya = fmt == "graya8p" and name != "strokes4k"
    rgba = fmt == "rgba8p" or name == "strokes4k" or graya  # "rgba": a read-modify-write (SrcOver) format
    bpp = 2 if graya else (4 if rgba else 1)
    rmw_color = (120, 255, 0, 0) if graya else (200, 120, 40, 255)  # opaque colour (gray, alpha) / (r, g, b, a)
    fmt_name = "Graya8p" if graya else ("Rgba8p" if rgba else "Matte8")
    wl = make_workload(name, batch, D.rank)
    if name == "strokes4k":
        def product_outline(path):
            p = fb.Plotter(fb.Raster(8, 8, Format.Rgba8p), device=D.local_rank)
            p.set_join(fb.JoinStyle.Round)
            return p.debug_stroke_ops(path)
        set_outlines(wl, batch, product_outline)
    if rgba:
        wl["ofmt"], wl["color"] = (1 if graya else 2), (rmw_color[:2] if graya else rmw_color)
    W, H = wl.get("w", wl["size"]), wl.get("h", wl["size"])
    config = {"workload": wl["desc"].replace("Matte8", fmt_name), "fills_per_step_per_gpu": batch, "raster": "%dx%d" % (W, H), "format": fmt_name, "pixels": PIXELS_NOTE,
              "l2": "each step writes fills_per_step rasters (>= 1 GiB per GPU) - outputs far exceed the 126 MB L2; inputs are a few KB"}
    b = Batch(W, H, Format.Graya8p if graya else (Format.Rgba8p if rgba else Format.Matte8), batch, device=D.local_rank)
    colors = np.tile(np.array(rmw_color, dtype=np.uint8), (batch, 1)) if rgba else None
    stream = torch.cuda.ExternalStream(b.stream(), device=D.local_rank)
    raster_bytes = W * H * bpp
    if name == "strokes4k":  # config 3 draws over an opaque (64,128,64,255) raster (examples/stroke2.rs:20-21)
        rect = fb.Path2D().absolute().move_to(0, 0).line_to(W, 0).line_to(W, H).line_to(0, H).close().finish()
        b.fill(np.tile(rect, batch), np.arange(batch + 1, dtype=np.uint64) * np.uint64(len(rect)),
               colors=np.tile(np.array([64, 128, 64, 255], dtype=np.uint8), (batch, 1))).sync()
        b.set_join(fb.JoinStyle.Round)

    # ---- value: ops resident in HBM ----
    b.upload(wl["ops"], wl["offs"], rules=wl["rules"], transforms=wl["tr"], colors=colors)
    for _ in range(warmup):
        b.run()
    D.barrier(b)
    # pixels per step by the reference's dense-row convention, from the product's own probe (top row of every job)
    tops = b.top_rows(0, batch).astype(np.int64)
    px_step = float(np.sum(np.where(tops == np.iinfo(np.int32).max, 0, W * np.maximum(0, H - np.maximum(tops, 0)))))
    sampler = ClockSampler(D.local_rank).start()
    fb.set_profiling(True)
    fb.tile_kernel_time(reset=True)
    l0 = fb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        b.run()
    ev1.record(stream)
    D.barrier(b)
    launches = fb.launch_count() - l0
    ms = D.max(ev0.elapsed_time(ev1))
    tile_ms, tile_n = fb.tile_kernel_time(reset=True)
    fb.set_profiling(False)
    clocks = sampler.stop()
    line = {"metric": "Gpx/s composited (%s %s)" % (LABEL[name], fmt_name), "value": px_step * D.world * steps / (ms * 1e-3) / 1e9, "unit": "Gpx/s", "n_gpus": D.world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic", "config": config, "paths_per_s": batch * D.world * steps / (ms * 1e-3), "clocks": clocks, "gpu_launches": int(launches)}
    if "outline_ms" in wl:
        config["stroke_outline_ms_total_host"] = wl["outline_ms"]
    # ---- roofline of the tile kernels (pixel term: 1 B/px store for Matte8, read + write for the SrcOver formats) ----
    peak, peak_kind = peaks()
    if tile_n:
        per_launch_ms = tile_ms / tile_n
        bytes_px = 2 * bpp if rgba else 1
        achieved = px_step * bytes_px / (per_launch_ms * 1e-3) / 1e9
        traffic, tsrc = recorded_traffic(name + ("_graya8p" if graya else ("_rgba8p" if rgba else "")))
        roof = {"kernel": "raster_tiles + raster_bins", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes_per_launch": px_step * bytes_px, "bytes_per_px": bytes_px,
                "avg_launch_ms": per_launch_ms, "launches_timed": tile_n, "share_of_step": tile_ms / ms if ms else None}
        if traffic:
            roof["traffic_frac"] = traffic / (per_launch_ms * 1e-3) / 1e9 / peak  # recorded DRAM bytes over this run's launch time
        if rgba:
            roof["note"] = ("read + write of every pixel (8 B/px Rgba8p, 4 B/px Graya8p) is what the reference's per-pixel SrcOver loop moves; here opaque spans are written unread and alpha-0 "
                            "spans write back only the pixels that change, so real DRAM traffic is lower and frac may exceed 1 - traffic_frac is the measured share of the peak")
        line["roofline"] = roof
    if not full:
        line["tile_ms_per_launch"] = tile_ms / max(tile_n, 1)
        del b
        return line

    # ---- e2e: host ops in, rasters out to pinned host memory, every step ----
    pinned = HostBuffer(batch * raster_bytes)
    e2e_steps = max(3, min(steps, 10))
    if name == "strokes4k":
        sops, soffs, _ = wl["stroke"]

        def step():  # the public call: outlines made inside the timed region (host stroker on threads), one device pass
            b.stroke(sops, soffs, colors=colors)
            b.read_into(0, batch, pinned.data_ptr(), pinned.numel())
        h2d = int(sops.nbytes + batch * 64)
    else:
        def step():
            b.fill(wl["ops"], wl["offs"], rules=wl["rules"], transforms=wl["tr"], colors=colors)
            b.read_into(0, batch, pinned.data_ptr(), pinned.numel())
        h2d = int(wl["ops"].nbytes + batch * 64)
    for _ in range(2):
        step()
    D.barrier(b)
    fb.transfer_bytes(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        step()
    e1.record(stream)
    D.barrier(b)
    e2e_ms = D.max(e0.elapsed_time(e1))
    wire_h2d, wire_d2h = fb.transfer_bytes(reset=True)
    line["e2e"] = {"value": px_step * D.world * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "Gpx/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(batch * raster_bytes),
                   "wire_h2d_bytes_per_step": wire_h2d // e2e_steps, "wire_d2h_bytes_per_step": wire_d2h // e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                   "note": "d2h_bytes_per_step is the logical size of the rasters handed back; the read-back travels packed (uniform 32-byte blocks as one code byte), wire_d2h is what crossed PCIe"}

    pinned.close()

    # ---- config 3 only: the batched public call alone (no read-back), stroker on the device against stroker on host threads ----
    if name == "strokes4k" and D.rank == 0:
        sops, soffs, _ = wl["stroke"]
        bs = {}
        keep_env = os.environ.get("FTL_DEVICE_STROKE")
        for mode, label in (("1", "device_stroker"), ("0", "host_stroker")):
            os.environ["FTL_DEVICE_STROKE"] = mode
            for _ in range(3):
                b.stroke(sops, soffs, colors=colors)
            b.sync()
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                b.stroke(sops, soffs, colors=colors)
            b.sync()
            dt = (time.perf_counter() - t0) / reps
            bs[label] = {"ms_per_step": 1e3 * dt, "strokes_per_s": batch / dt, "value": px_step / dt / 1e9, "unit": "Gpx/s"}
        if keep_env is None:
            os.environ.pop("FTL_DEVICE_STROKE", None)
        else:
            os.environ["FTL_DEVICE_STROKE"] = keep_env
        bs["what"] = ("ftl_batch_stroke of %d strokes + ftl_batch_sync, rasters stay in HBM, host wall clock: flatten with widths, outline (stroker.rs:204-416) and fill on the "
                      "device (stroke_kernels.cuh), against the same call with the outlines made by the host stroker on threads" % batch)
        line["batch_stroke"] = bs

    # ---- config 4 only: the same paths STROKED (pen width 3, Round joins): the stroker on the device against host threads ----
    if name == "batch512" and D.rank == 0 and fmt == "matte8":
        n_jobs = len(wl["offs"]) - 1
        pw = np.zeros(1, dtype=wl["ops"].dtype)
        pw["tag"] = 5
        pw["v"][0, 0] = 3.0
        sops = np.insert(wl["ops"], np.asarray(wl["offs"][:-1], dtype=np.int64), pw)
        soffs = np.asarray(wl["offs"], dtype=np.uint64) + np.arange(n_jobs + 1, dtype=np.uint64)
        b.set_join(fb.JoinStyle.Round)
        bs = {}
        keep_env = os.environ.get("FTL_DEVICE_STROKE")
        sums = {}
        for mode, label in (("1", "device_stroker"), ("0", "host_stroker")):
            os.environ["FTL_DEVICE_STROKE"] = mode
            for _ in range(2):
                b.stroke(sops, soffs, transforms=wl["tr"])
            b.sync()
            sums[label] = b.checksums(0, n_jobs).copy()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                b.stroke(sops, soffs, transforms=wl["tr"])
            b.sync()
            dt = (time.perf_counter() - t0) / reps
            bs[label] = {"ms_per_step": 1e3 * dt, "strokes_per_s": n_jobs / dt}
        if keep_env is None:
            os.environ.pop("FTL_DEVICE_STROKE", None)
        else:
            os.environ["FTL_DEVICE_STROKE"] = keep_env
        bs["same_pixels"] = bool(np.array_equal(sums["device_stroker"], sums["host_stroker"]))
        bs["what"] = ("ftl_batch_stroke of the %d paths of this workload (pen width 3, Round joins) + ftl_batch_sync, rasters stay in HBM, host wall clock: stroker on the "
                      "device (stroke_kernels.cuh) against the host stroker on 8 threads; same_pixels compares the raster checksums of the two" % n_jobs)
        line["batch_stroke"] = bs

    # ---- config 3 only: the single-plotter call, Plotter::stroke (host flatten + host stroker -> device fill) ----
    if name == "strokes4k" and D.rank == 0:
        from footile_b200 import scenes as _scenes
        paths = list(_scenes.stroke_scenes(30.0).values())
        bg = np.tile(np.array([64, 128, 64, 255], dtype=np.uint8), (H, W))
        plotters = []
        for _ in paths:
            pl = fb.Plotter(fb.Raster(W, H, Format.Rgba8p, bg), device=D.local_rank)
            pl.set_join(fb.JoinStyle.Round)
            plotters.append(pl)
        for pl, path in zip(plotters, paths):  # warm-up
            pl.stroke(path, (255, 255, 0, 255))
            pl.sync()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            for pl, path in zip(plotters, paths):
                pl.stroke(path, (255, 255, 0, 255))
            for pl in plotters:
                pl.sync()
        dt = time.perf_counter() - t0
        n_calls = reps * len(paths)
        line["plotter_stroke"] = {"calls": n_calls, "ms_per_stroke": 1e3 * dt / n_calls, "strokes_per_s": n_calls / dt, "value": px_step / batch * n_calls / dt / 1e9, "unit": "Gpx/s",
                                  "what": "ftl_stroke through the Plotter mirror, one call per scene, rasters stay in HBM: the library's default placement of the stroker "
                                          "(host below 512 ops: these scenes have 3-12 ops each), one device fill"}
        del plotters

    # ---- cpu baseline: oracle, one thread, bounded sample (rank 0, N=1 only) ----
    line["cpu_baseline"] = None
    if D.rank == 0 and D.world == 1:
        n_cpu = min(args.cpu_fills or {"heptagram": 8, "batch512": 256, "fishy256": 4096, "strokes4k": 6}[name], batch)
        jobs = list(range(n_cpu))
        cpu_run(wl, jobs[: max(1, n_cpu // 4)], 1)
        reps, t = 0, 0.0
        while t < 8.0 and reps < 50:
            t += cpu_run(wl, jobs, 1)
            reps += 1
        line["cpu_baseline"] = {"value": px_step / batch * n_cpu * reps / t / 1e9, "unit": "Gpx/s", "cores": 1, "kind": "port",
                                "sample": "%d fills x %d repeats of this workload, single thread, SSSE3 accumulate, rasters pre-allocated, timed inside the oracle" % (n_cpu, reps),
                                "paths_per_s": n_cpu * reps / t}
    del b
    return line


def default_batch(name, fmt):
    rmw = fmt != "matte8" or name == "strokes4k"
    return {"heptagram": 64 if rmw else 256, "batch512": 1024 if rmw else 4096, "fishy256": 16384, "strokes4k": 36}[name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="heptagram", choices=WORKLOADS)
    ap.add_argument("--batch", type=int, default=0, help="fills per step per GPU (default 256 heptagram / 4096 batch512)")
    ap.add_argument("--cpu-fills", type=int, default=0, help="fills in the cpu_baseline sample (default: sized for ~10 s)")
    ap.add_argument("--format", default="matte8", choices=["matte8", "rgba8p", "graya8p"], help="pixel format of the rasters")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e, cpu_baseline and secondary legs (for runs under ncu)")
    ap.add_argument("--no-secondary", action="store_true", help="default workload: skip the config 4 / config 5 / latency figures")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))

    if args.workload == "bigraster":
        return run_bigraster(args)

    if args.workload == "latency":
        if rank != 0:
            return
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "latency is reported by the `ours` arm beside its own cpu figures"}))
            return
        D = Dist()
        print(json.dumps({"metric": "microseconds per Plotter::fill call (benches/fishyb.rs)", "unit": "us", "higher_is_better": False, "n_gpus": 1,
                          "data": "synthetic", "dtype": DTYPE, "latency": latency_numbers(D.local_rank)}))
        return

    batch = args.batch or default_batch(args.workload, args.format)

    # ---------------- reference arm: the oracle on all host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        graya = args.format == "graya8p" and args.workload != "strokes4k"
        rgba = args.format == "rgba8p" or args.workload == "strokes4k" or graya
        fmt_name = "Graya8p" if graya else ("Rgba8p" if rgba else "Matte8")
        wl = make_workload(args.workload, batch, 0)
        if args.workload == "strokes4k":
            def oracle_outline(path):
                import oracle
                o = oracle.Plotter(8, 8, oracle.RGBA8P)
                o.set_join(oracle.ROUND, 0.0)
                return o.debug_stroke_ops(path)
            set_outlines(wl, batch, oracle_outline)
        if rgba:
            wl["ofmt"], wl["color"] = (1 if graya else 2), ((120, 255) if graya else (200, 120, 40, 255))
        config = {"workload": wl["desc"].replace("Matte8", fmt_name), "fills_per_step_per_gpu": batch, "raster": "%dx%d" % (wl.get("w", wl["size"]), wl.get("h", wl["size"])),
                  "format": fmt_name, "pixels": PIXELS_NOTE,
                  "l2": "each step writes fills_per_step rasters (>= 1 GiB per GPU) - outputs far exceed the 126 MB L2; inputs are a few KB"}
        cores = os.cpu_count() or 1
        per_step = max(cores, min(batch, cores * {"heptagram": 2, "batch512": 64, "fishy256": 1024, "strokes4k": 1}[args.workload]))
        jobs = list(range(per_step))
        px = sum(oracle_pixels(wl, jobs))
        for _ in range(min(args.warmup, 2)):
            cpu_run(wl, jobs, cores)
        t = 0.0
        for _ in range(args.steps):
            t += cpu_run(wl, jobs, cores)
        v = px * args.steps / t / 1e9
        sample = "%d fills per step on %d C++ threads (one Plotter per fill, rasters pre-allocated, timed inside the oracle)" % (per_step, cores)
        print(json.dumps({"impl": "reference", "metric": "Gpx/s composited (%s %s)" % (LABEL[args.workload], fmt_name), "value": v, "unit": "Gpx/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "Gpx/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "paths_per_s": per_step * args.steps / t,
                          "note": "CPU restatement of footile (the Rust reference cannot be built in this image)"}))
        return

    # ---------------- our arm ----------------
    D = Dist()
    line = run_batch(D, args, args.workload, batch, args.format, args.steps, args.warmup, full=not args.kernel_only)
    if args.kernel_only:
        line["kernel_only"] = True
    elif args.workload == "heptagram" and args.format == "matte8" and not args.no_secondary:
        # the two multi-GPU configs of BASELINE.json at this N, and the reference's own benchmark (latency), beside the headline
        sec = {}
        try:
            r4 = run_batch(D, args, "batch512", default_batch("batch512", "matte8"), "matte8", 10, 3, full=False)
            sec["batch512_paths_per_s"] = r4["paths_per_s"]
            sec["batch512_ms_per_step"] = r4["ms_per_step"]
            sec["batch512_gpx_per_s"] = r4["value"]
            sec["batch512_roofline_frac"] = (r4.get("roofline") or {}).get("frac")
            sec["batch512_what"] = "config 4: 4096 random 64-curve paths x 512^2 Matte8 per GPU per step, ops resident (weak scaling: per-GPU work fixed)"
            r5 = bigraster_ours(D, 3, 3, full=False)
            sec["bigraster_ms_per_fill"] = r5["ms_per_fill"]
            sec["bigraster_gpx_per_s"] = r5["value"]
            sec["bigraster_tile_ms_per_launch"] = r5["tile_ms_per_launch"]
            sec["bigraster_what"] = "config 5: one 32768^2 Matte8 raster, 10.24 M edges, one EvenOdd fill split into %d row bands, ops resident (strong scaling: total work fixed; max over ranks)" % D.world
            if D.rank == 0:
                sec["latency"] = latency_numbers(D.local_rank, cpu=D.world == 1)
        except Exception as e:  # the headline stands on its own
            sec["error"] = repr(e)
        line["secondary"] = sec
    if D.rank == 0:
        print(json.dumps(line))
    D.done()


if __name__ == "__main__":
    main()
