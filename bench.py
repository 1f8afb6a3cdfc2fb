#!/usr/bin/env python
"""bench.py — throughput of the footile hot path on B200 (and the CPU baseline beside it).

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload heptagram|batch512]
A step is one pass of the hot path over one batch of synthetic paths.  Default workload =
BASELINE.json configs[1]: heptagram fills (NonZero / EvenOdd alternating) into 4096x4096 Matte8
rasters, `--batch` rasters per step (default 256 = 4 GiB of output, far beyond the 126 MB L2).
  value : Gpx/s with the path ops already resident in HBM (ftl_batch_run), CUDA events on the
          library's stream, max over ranks.
  e2e   : the same metric through the C ABI with HOST buffers: ftl_batch_fill(host ops) then
          ftl_batch_read of every raster into pinned host memory, inside the timed region.
  roofline : raster_tiles kernel, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json.
  cpu_baseline : the oracle (CPU restatement of footile, 1 thread) on a bounded sample, rank 0, N=1.
--impl reference times the oracle on all host threads (one plotter per thread).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PIXELS_NOTE = "px = W * (H - max(top_row,0)) per fill: the reference resolves every pixel of those rows (fig.rs:497,539)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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

    def _sample(self):
        self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def _loop(self):
        while self.run:
            try:
                self._sample()
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
        self.huge = False
        if os.environ.get("FTL_BENCH_NO_THP") != "1":
            try:
                self.m.madvise(mmap.MADV_HUGEPAGE)
                self.huge = True
            except Exception:
                pass
        self.a = np.frombuffer(self.m, dtype=np.uint8)
        self.a[:] = 0  # fault the pages in
        self.registered = False
        try:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.a.ctypes.data, self.n, 0)
            self.registered = int(rc) == 0
        except Exception:
            pass

    def data_ptr(self):
        return self.a.ctypes.data

    def numel(self):
        return self.n


def host_result_buffer(nbytes):
    return HostBuffer(nbytes)


# ---- workloads -----------------------------------------------------------------
def make_workload(name, batch, rank, outline_of=None):
    from footile_b200 import scenes
    if name == "heptagram":
        size = 4096
        path = scenes.heptagram_abs()
        ops = np.tile(path, batch)
        offs = np.arange(batch + 1, dtype=np.uint64) * np.uint64(len(path))
        rules = (np.arange(batch) & 1).astype(np.uint8)
        tr = np.tile(scenes.heptagram_transform(size), (batch, 1))
        return dict(size=size, ops=ops, offs=offs, rules=rules, tr=tr, desc="heptagram {7/2} fill, NonZero/EvenOdd alternating, 4096x4096 Matte8")
    if name == "batch512":
        size = 512
        ops, offs, rules = scenes.random_curve_paths(rank * batch, batch)
        return dict(size=size, ops=ops, offs=offs, rules=rules, tr=None, desc="random quad/cubic paths (64 segments), 512x512 Matte8 per path")
    if name == "fishy256":
        size = 256
        path = scenes.fishy_bench()
        ops = np.tile(path, batch)
        offs = np.arange(batch + 1, dtype=np.uint64) * np.uint64(len(path))
        rules = np.zeros(batch, dtype=np.uint8)
        tr = np.tile(np.array([2, 0, 0, 0, 2, 0], dtype=np.float32), (batch, 1))
        return dict(size=size, ops=ops, offs=offs, rules=rules, tr=tr, desc="benches/fishyb.rs fill_256: fishy path, scale(2,2), 256x256 Matte8 per fill")
    if name == "strokes4k":
        # config 3: the six stroke scenes x30 with Round joins; the OUTLINES (what Plotter::stroke hands to fill,
        # plotter.rs:361-364) are made once, outside the timed region, by `outline_of` (the product's device flatten +
        # host stroker in our arm, the oracle's in the reference arm) and then filled NonZero into 3840x2160 Rgba8p.
        t0 = time.perf_counter()
        outs = [outline_of(p) for p in scenes.stroke_scenes(30.0).values()]
        t_outline = time.perf_counter() - t0
        parts = [outs[j % len(outs)] for j in range(batch)]
        offs = np.zeros(batch + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(p) for p in parts])
        return dict(size=0, w=3840, h=2160, ops=np.concatenate(parts), offs=offs, rules=np.zeros(batch, dtype=np.uint8), tr=None,
                    outline_ms=1e3 * t_outline, desc="stroke.rs/stroke2.rs/round.rs/over.rs/teeth.rs/curve.rs x30, Round joins, butt ends: outline fill into 3840x2160 Rgba8p")
    raise SystemExit("unknown workload " + name)


def oracle_pixels(wl, sample):
    """Pixels per fill by the reference's dense-row convention, from the oracle's top_row."""
    import oracle
    px = []
    W, H = wl.get("w", wl["size"]), wl.get("h", wl["size"])
    for j in sample:
        o = oracle.Plotter(8, 8, oracle.MATTE8)  # tiny raster: only (dir, top_row) are needed
        if wl["tr"] is not None:
            o.set_transform(wl["tr"][j])
        o.fill(int(wl["rules"][j]), wl["ops"][int(wl["offs"][j]): int(wl["offs"][j + 1])], (255,))
        top = o.last_info()["top_row"]
        px.append(W * max(0, H - max(top, 0)))
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


def bigraster(args, rank, local_rank, world):
    """BASELINE.json configs[4]: ONE 32768x32768 Matte8 raster, 160 000 closed 64-gons (10.24 M edges) in one fill,
    rows split into `world` bands, one band per GPU, no collective (every rank gets the same ops and recomputes
    (dir, top_row)).  Strong scaling: the work is fixed, `value` = pixels of the whole raster / slowest rank's time.
    The reference arm times the order-free oracle on a bounded sample of 16-row stripes, one per host thread."""
    from footile_b200 import scenes
    size, polys = 32768, 160000
    metric, unit = "Gpx/s composited (one 32768^2 Matte8 raster, 10.24 M edges, row bands)", "Gpx/s"
    config = {"workload": "160 000 closed 64-gon sub-figures (seed 0xB160000^i) in ONE EvenOdd fill of a 32768x32768 Matte8 raster",
              "raster": "32768x32768", "format": "Matte8", "bands": world, "pixels": PIXELS_NOTE,
              "l2": "1 GiB raster and 164 MB of edges per fill: far beyond the 126 MB L2"}
    ops = scenes.random_polygons(0, polys, vertices=64, size=size, extent=2048)
    if args.impl == "reference":
        if rank != 0:
            return
        import oracle
        cores = os.cpu_count() or 1
        rows_each = 16
        rng = np.random.default_rng(5)
        starts = [int(r) for r in rng.integers(0, size - rows_each, cores)]

        def stripe(r0):
            o = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True)
            o.set_rows(r0, r0 + rows_each)
            o.fill(1, ops, (255,))

        ts = []
        for _ in range(max(1, min(args.steps, 3))):
            t0 = time.perf_counter()
            ths = [threading.Thread(target=stripe, args=(r0,)) for r0 in starts]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            ts.append(time.perf_counter() - t0)
        t = min(ts)
        v = size * rows_each * cores / t / 1e9
        sample = "%d stripes of %d rows (order-free oracle, u32 vertex ids), one per thread on %d threads; every stripe walks all 10.24 M edges" % (cores, rows_each, cores)
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": len(ts), "warmup": 0,
                          "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32 fixed-point 16.16 / u8",
                          "data": "synthetic", "config": config, "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    import torch.distributed as dist
    import footile_b200 as fb
    from footile_b200 import Format, sharding
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r0, r1 = sharding.band_rows(size, rank, world)

    p = fb.Plotter.with_clear(size, size, Format.Matte8, device=local_rank, rows=(r0, r1))
    for _ in range(max(1, min(args.warmup, 3))):
        p.fill(1, ops, (255,))
    p.sync()
    if world > 1:
        dist.barrier()
    fb.set_profiling(True)
    fb.tile_kernel_time(reset=True)
    l0 = fb.launch_count()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        p.fill(1, ops, (255,))
    p.sync()
    dt = sharding.max_over_ranks(time.perf_counter() - t0, device="cuda" if world > 1 else None)
    launches = fb.launch_count() - l0
    tile_ms, tile_n = fb.tile_kernel_time(reset=True)
    fb.set_profiling(False)
    top = p.debug_last_fill()["top_row"]
    px = size * (size - max(top, 0))
    value = px * steps / dt / 1e9
    pinned = torch.empty((r1 - r0) * size, dtype=torch.uint8, pin_memory=True)
    t0 = time.perf_counter()
    for _ in range(2):
        p.fill(1, ops, (255,))
        fb._lib.check(fb._lib.lib().ftl_read_raster(p._handle, pinned.data_ptr(), pinned.numel()))
    e2e_dt = sharding.max_over_ranks((time.perf_counter() - t0) / 2, device="cuda" if world > 1 else None)
    peak, peak_kind = peaks()
    if rank == 0:
        per_launch_ms = tile_ms / max(tile_n, 1)
        ach = px / world / (per_launch_ms * 1e-3) / 1e9
        print(json.dumps({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32 fixed-point 16.16 / u8", "data": "synthetic",
                          "config": config, "edges_per_s": polys * 64 * steps / dt,
                          "timing": "host clock around fill+sync, max over ranks: every step includes the H2D copy of 291 MB of ops (the Plotter API re-sends them)",
                          "e2e": {"value": px / e2e_dt / 1e9, "unit": unit, "h2d_bytes_per_step": int(ops.nbytes), "d2h_bytes_per_step": int((r1 - r0) * size), "ms_per_step": 1e3 * e2e_dt},
                          "gpu_launches": int(launches),
                          "roofline": {"kernel": "raster_tiles", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_kind": peak_kind,
                                       "traffic": None, "avg_launch_ms": per_launch_ms, "note": "scatter/issue-bound: ~200 (edge,row) items per edge"},
                          "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="heptagram", choices=["heptagram", "batch512", "fishy256", "strokes4k", "bigraster"])
    ap.add_argument("--batch", type=int, default=0, help="fills per step per GPU (default 256 heptagram / 4096 batch512)")
    ap.add_argument("--cpu-fills", type=int, default=0, help="fills in the cpu_baseline sample (default: sized for ~10 s)")
    ap.add_argument("--format", default="matte8", choices=["matte8", "rgba8p", "graya8p"], help="pixel format of the rasters")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and cpu_baseline legs (for runs under ncu)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    graya = args.format == "graya8p" and args.workload != "strokes4k"
    rgba = args.format == "rgba8p" or args.workload == "strokes4k" or graya  # "rgba": a read-modify-write (SrcOver) format
    bpp = 2 if graya else (4 if rgba else 1)
    ofmt_id = 1 if graya else 2  # oracle.GRAYA8P / oracle.RGBA8P
    rmw_color = (120, 255, 0, 0) if graya else (200, 120, 40, 255)  # opaque colour (gray, alpha) / (r, g, b, a)
    batch = args.batch or ((64 if rgba else 256) if args.workload == "heptagram" else {"batch512": 1024 if rgba else 4096, "fishy256": 16384, "strokes4k": 36, "bigraster": 1}[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    unit = "Gpx/s"
    fmt_name = "Graya8p" if graya else ("Rgba8p" if rgba else "Matte8")
    metric = "Gpx/s composited (%s %s)" % ({"heptagram": "heptagram fill 4096^2", "batch512": "100k-path batch 512^2", "fishy256": "fishyb fill_256 batch 256^2", "strokes4k": "stroke scenes x30 3840x2160", "bigraster": "one 32768^2 raster"}[args.workload], fmt_name)
    config = {"workload": None, "fills_per_step_per_gpu": batch, "raster": None, "format": fmt_name, "pixels": PIXELS_NOTE,
              "l2": "each step writes fills_per_step rasters (>= 1 GiB per GPU) - outputs far exceed the 126 MB L2; inputs are a few KB"}

    # ---------------- config 5: one huge raster split into row bands (strong scaling) ----------------
    if args.workload == "bigraster":
        return bigraster(args, rank, local_rank, world)

    # ---------------- reference arm: the oracle on all host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        def oracle_outline(path):
            import oracle
            o = oracle.Plotter(8, 8, oracle.RGBA8P)
            o.set_join(oracle.ROUND, 0.0)
            return o.debug_stroke_ops(path)

        wl = make_workload(args.workload, batch, 0, oracle_outline)
        if rgba:
            wl["ofmt"], wl["color"] = ofmt_id, rmw_color[:2] if graya else rmw_color
        config["workload"] = wl["desc"].replace("Matte8", fmt_name)
        config["raster"] = "%dx%d" % (wl.get("w", wl["size"]), wl.get("h", wl["size"]))
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
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "i32 fixed-point 16.16 / u8", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "paths_per_s": per_step * args.steps / t,
                          "note": "CPU restatement of footile (the Rust reference cannot be built in this image)"}))
        return

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist
    import footile_b200 as fb
    from footile_b200 import Batch, Format

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    def product_outline(path):
        p = fb.Plotter(fb.Raster(8, 8, Format.Rgba8p), device=local_rank)
        p.set_join(fb.JoinStyle.Round)
        return p.debug_stroke_ops(path)  # device flatten + host stroker

    wl = make_workload(args.workload, batch, rank, product_outline)
    if rgba:
        wl["ofmt"], wl["color"] = ofmt_id, rmw_color[:2] if graya else rmw_color
    W, H = wl.get("w", wl["size"]), wl.get("h", wl["size"])
    config["workload"] = wl["desc"].replace("Matte8", fmt_name)
    config["raster"] = "%dx%d" % (W, H)
    if "outline_ms" in wl:
        config["stroke_outline_ms_total_host"] = wl["outline_ms"]
    b = Batch(W, H, Format.Graya8p if graya else (Format.Rgba8p if rgba else Format.Matte8), batch, device=local_rank)
    colors = np.tile(np.array(rmw_color, dtype=np.uint8), (batch, 1)) if rgba else None
    stream = torch.cuda.ExternalStream(b.stream(), device=local_rank)
    px_fill = oracle_pixels(wl, range(batch) if args.workload == "batch512" else range(min(batch, 2)))
    if len(px_fill) == batch:
        px_step = float(sum(px_fill))
    else:  # heptagram: every fill has the same top row
        px_step = float(px_fill[0]) * batch
    raster_bytes = W * H * bpp
    if args.workload == "strokes4k":  # config 3 draws over an opaque (64,128,64,255) raster (examples/stroke2.rs:20-21)
        rect = fb.Path2D().absolute().move_to(0, 0).line_to(W, 0).line_to(W, H).line_to(0, H).close().finish()
        b.fill(np.tile(rect, batch), np.arange(batch + 1, dtype=np.uint64) * np.uint64(len(rect)),
               colors=np.tile(np.array([64, 128, 64, 255], dtype=np.uint8), (batch, 1))).sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        b.sync()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: ops resident in HBM ----
    b.upload(wl["ops"], wl["offs"], rules=wl["rules"], transforms=wl["tr"], colors=colors)
    for _ in range(args.warmup):
        b.run()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    fb.set_profiling(True)
    fb.tile_kernel_time(reset=True)
    l0 = fb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        b.run()
    ev1.record(stream)
    barrier()
    launches = fb.launch_count() - l0
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    tile_ms, tile_n = fb.tile_kernel_time(reset=True)
    fb.set_profiling(False)
    clocks = sampler.stop()
    value = px_step * world * args.steps / (ms * 1e-3) / 1e9
    paths_per_s = batch * world * args.steps / (ms * 1e-3)

    # ---- e2e: host ops in, rasters out to pinned host memory, every step ----
    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"metric": metric, "value": value, "unit": unit, "ms_per_step": ms / args.steps, "kernel_only": True,
                              "tile_ms_per_launch": tile_ms / max(tile_n, 1), "gpu_launches": int(launches)}))
        return
    pinned = host_result_buffer(batch * raster_bytes)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        b.fill(wl["ops"], wl["offs"], rules=wl["rules"], transforms=wl["tr"], colors=colors)
        b.read_into(0, batch, pinned.data_ptr(), pinned.numel())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        b.fill(wl["ops"], wl["offs"], rules=wl["rules"], transforms=wl["tr"], colors=colors)
        b.read_into(0, batch, pinned.data_ptr(), pinned.numel())
    e1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_val = px_step * world * e2e_steps / (e2e_ms * 1e-3) / 1e9
    h2d = int(wl["ops"].nbytes + batch * 64)
    d2h = int(batch * raster_bytes)

    # ---- config 3 only: the public call itself, Plotter::stroke (device flatten -> host stroker -> device fill) ----
    api_stroke = None
    if args.workload == "strokes4k" and rank == 0:
        from footile_b200 import scenes as _scenes
        paths = list(_scenes.stroke_scenes(30.0).values())
        bg = np.tile(np.array([64, 128, 64, 255], dtype=np.uint8), (H, W))
        plotters = []
        for _ in paths:
            pl = fb.Plotter(fb.Raster(W, H, Format.Rgba8p, bg), device=local_rank)
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
        api_stroke = {"calls": n_calls, "ms_per_stroke": 1e3 * dt / n_calls, "strokes_per_s": n_calls / dt,
                      "value": px_step / batch * n_calls / dt / 1e9, "unit": unit,
                      "what": "ftl_stroke through the Plotter mirror, one call per scene, rasters stay in HBM: device flatten, D2H, host stroker, H2D, device fill"}
        del plotters

    # ---- roofline of the tile kernel (pixel term: 1 B/px store for Matte8) ----
    peak, peak_kind = peaks()
    roof = None
    if tile_n:
        per_launch_ms = tile_ms / tile_n
        achieved = px_step * (2 * bpp if rgba else 1) / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.workload + ("_graya8p" if graya else ("_rgba8p" if rgba else "")))
        except Exception:
            pass
        roof = {"kernel": "raster_tiles", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_kind": peak_kind, "traffic": traffic, "algorithmic_bytes_per_launch": px_step * (2 * bpp if rgba else 1), "bytes_per_px": 2 * bpp if rgba else 1, "avg_launch_ms": per_launch_ms,
                "launches_timed": tile_n, "share_of_step": tile_ms / ms if ms else None}
        if traffic:
            roof["traffic_frac"] = traffic / (per_launch_ms * 1e-3) / 1e9 / peak  # measured DRAM bytes (ncu) over this run's launch time
        if rgba:
            roof["note"] = ("read + write of every pixel (8 B/px Rgba8p, 4 B/px Graya8p) is what the reference's per-pixel SrcOver loop moves; here opaque spans are written unread and alpha-0 "
                            "spans write back only the pixels that change, so real DRAM traffic is lower and frac may exceed 1 - traffic_frac is the measured share of the peak")

    # ---- cpu baseline: oracle, one thread, bounded sample (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1:
        n_cpu = args.cpu_fills or {"heptagram": 8, "batch512": 256, "fishy256": 4096, "strokes4k": 6}[args.workload]
        n_cpu = min(n_cpu, batch)
        jobs = list(range(n_cpu))
        cpu_run(wl, jobs[: max(1, n_cpu // 4)], 1)
        reps, t = 0, 0.0
        while t < 8.0 and reps < 50:
            t += cpu_run(wl, jobs, 1)
            reps += 1
        cpx = px_step / batch * n_cpu
        cpu = {"value": cpx * reps / t / 1e9, "unit": unit, "cores": 1, "kind": "port",
               "sample": "%d fills x %d repeats of this workload, single thread, SSSE3 accumulate, rasters pre-allocated, timed inside the oracle" % (n_cpu, reps),
               "paths_per_s": n_cpu * reps / t}

    if rank == 0:
        print(json.dumps({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "i32 fixed-point 16.16 / u8", "data": "synthetic", "config": config, "paths_per_s": paths_per_s,
                          "clocks": clocks, "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
                          "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                          **({"plotter_stroke": api_stroke} if api_stroke else {})}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
