"""Scene / workload definitions of the five BASELINE.json configs (geometry only).

Host-side numpy; shared by tests/ and bench.py so that the GPU path, the CPU
oracle and the benchmark all draw exactly the same ops.  The example scenes
follow the reference's examples/*.rs and benches/fishyb.rs (cited per
function); the two synthetic configs use a counter-based SplitMix64 so that
sharding over GPUs never changes a path.
"""
import numpy as np

from .path import OP_DTYPE, OpTag, Path2D

_F32 = np.float32


# ---- config 1: fishy -------------------------------------------------------
def fishy_example():
    """examples/fishy.rs:9-27: (fish, eye) paths; 128x128 Rgba8p raster."""
    fish = (Path2D().relative().pen_width(3.0).move_to(112.0, 24.0).line_to(-32.0, 24.0)
            .cubic_to(-96.0, -48.0, -96.0, 80.0, 0.0, 32.0).line_to(32.0, 24.0).line_to(-16.0, -40.0).close().finish())
    eye = (Path2D().relative().pen_width(2.0).move_to(24.0, 48.0).line_to(8.0, 8.0).move_to(0.0, -8.0)
           .line_to(-8.0, 8.0).finish())
    return fish, eye


def fishy_bench():
    """benches/fishyb.rs:41-51: the bench fish (drawn with scale(2,2) on 16^2 / 256^2 Matte8)."""
    return (Path2D().relative().move_to(112.0, 16.0).line_to(-48.0, 32.0).cubic_to(-64.0, -48.0, -64.0, 80.0, 0.0, 32.0)
            .line_to(48.0, 32.0).line_to(-32.0, -48.0).close().finish())


# ---- config 2: heptagram ---------------------------------------------------
def heptagram():
    """examples/heptagram.rs:17-23: unit {7/2} star, theta_n = 4*pi*n/7 evaluated in f32."""
    pb = Path2D().move_to(_F32(np.cos(_F32(0))), _F32(np.sin(_F32(0))))
    # the builder is in relative mode by default (path.rs:59); the example's points are deltas of that mode
    for n in range(1, 7):
        th = _F32(_F32(_F32(np.pi) * _F32(4.0)) * _F32(n)) / _F32(7.0)
        pb = pb.line_to(_F32(np.cos(th)), _F32(np.sin(th)))
    return pb.close().finish()


def heptagram_abs():
    """The heptagram with absolute unit-circle vertices (the figure SURVEY §8d config 2 describes)."""
    pb = Path2D().absolute().move_to(_F32(1.0), _F32(0.0))
    for n in range(1, 7):
        th = _F32(_F32(_F32(np.pi) * _F32(4.0)) * _F32(n)) / _F32(7.0)
        pb = pb.line_to(_F32(np.cos(th)), _F32(np.sin(th)))
    return pb.close().finish()


def heptagram_transform(size):
    """Explicit transform of SURVEY §8d config 2: radius 0.45*size, centred."""
    s = _F32(0.45) * _F32(size)
    c = _F32(size) / _F32(2)
    return np.array([s, 0, c, 0, s, c], dtype=np.float32)


# ---- config 3: stroke scenes -----------------------------------------------
def stroke_scenes(scale=1.0):
    """examples/{stroke,stroke2,round,over,teeth,curve}.rs geometry, coordinates and pen widths
    multiplied by `scale` on the host (identity transform avoids the double-transform quirk)."""
    k = float(scale)

    def P():
        return Path2D().relative()

    s = {}
    s["stroke"] = (P().pen_width(5.0 * k).move_to(16.0 * k, 48.0 * k).line_to(32.0 * k, 0.0).line_to(-16.0 * k, -32.0 * k)
                   .close().finish())  # stroke.rs:9-16
    s["stroke2"] = (P().pen_width(6.0 * k).move_to(16.0 * k, 15.0 * k).line_to(32.0 * k, 1.0 * k).line_to(-32.0 * k, 1.0 * k)
                    .line_to(32.0 * k, 15.0 * k).line_to(-32.0 * k, 15.0 * k).line_to(32.0 * k, 1.0 * k)
                    .line_to(-32.0 * k, 1.0 * k).finish())  # stroke2.rs:9-19
    s["round"] = (P().pen_width(40.0 * k).move_to(10.0 * k, 60.0 * k).line_to(50.0 * k, 0.0).line_to(0.0, -50.0 * k)
                  .finish())  # round.rs:9-15
    s["over"] = (P().pen_width(8.0 * k).move_to(32.0 * k, 16.0 * k).line_to(16.0 * k, 16.0 * k).line_to(-16.0 * k, 16.0 * k)
                 .line_to(-16.0 * k, -16.0 * k).line_to(16.0 * k, -16.0 * k).line_to(0.0, 32.0 * k).finish())  # over.rs:9-18
    t = P().move_to(0.0, 8.0 * k)
    for i in range(8):
        t = t.line_to(8.0 * k, (8.0 if i % 2 == 0 else -8.0) * k)
    t = t.move_to(-64.0 * k, 32.0 * k)
    for i in range(8):
        t = t.line_to(8.0 * k, (8.0 if i % 2 == 0 else -8.0) * k)
    s["teeth"] = t.finish()  # teeth.rs:8-29 (default pen width 1 -> scaled by the caller if wanted)
    s["curve"] = (P().pen_width(0.0).move_to(64.0 * k, 48.0 * k).pen_width(18.0 * k)
                  .cubic_to(-64.0 * k, -48.0 * k, -64.0 * k, 80.0 * k, 0.0, 32.0 * k).finish())  # curve.rs:9-15
    return s


# ---- SplitMix64, counter based ----------------------------------------------
_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _draws(seeds, n):
    """n successive SplitMix64 outputs for every seed: uint64 [len(seeds), n]."""
    k = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        state = seeds[:, None] + k[None, :] * _GAMMA
    return _mix(state)


def _unit(u64):
    """Top 24 bits -> f32 in [0, 1)."""
    return (u64 >> np.uint64(40)).astype(np.float32) / _F32(16777216.0)


# ---- config 4: batch of random quad/cubic paths ------------------------------
def random_curve_paths(first, count, segments=64, size=512):
    """Paths first..first+count-1 of config 4.  Returns (ops, offsets u64[count+1], rules u8[count]).

    Path i: seed = splitmix64 stream of (0xF00711E5 ^ i); start uniform in [size/16, 15*size/16)^2;
    `segments` segments, each Quad or Cubic with probability 1/2; every control / end point is the
    previously generated point plus a uniform offset in [-3*size/16, 3*size/16)^2, clamped to
    [0, size); closed.  Rule NonZero for even i, EvenOdd for odd i.
    """
    idx = np.arange(first, first + count, dtype=np.uint64)
    seeds = _mix(np.uint64(0xF00711E5) ^ idx)
    per_seg = 7
    d = _draws(seeds, 2 + per_seg * segments)
    lo, span = _F32(size / 16.0), _F32(size * 14.0 / 16.0)
    off, ospan = _F32(-3.0 * size / 16.0), _F32(6.0 * size / 16.0)
    hi = np.nextafter(_F32(size), _F32(0))
    cur = np.stack([_unit(d[:, 0]) * span + lo, _unit(d[:, 1]) * span + lo], axis=1).astype(np.float32)
    n_ops = segments + 2
    ops = np.zeros((count, n_ops), dtype=OP_DTYPE)
    ops["tag"][:, 0] = OpTag.Move
    ops["v"][:, 0, 0:2] = cur
    for s in range(segments):
        base = 2 + per_seg * s
        cubic = (d[:, base] >> np.uint64(63)).astype(bool)
        pts = []
        for k in range(3):
            dx = _unit(d[:, base + 1 + 2 * k]) * ospan + off
            dy = _unit(d[:, base + 2 + 2 * k]) * ospan + off
            nxt = np.stack([cur[:, 0] + dx, cur[:, 1] + dy], axis=1).astype(np.float32)
            nxt = np.clip(nxt, _F32(0), hi)
            if k < 2:
                cur = nxt
            else:
                cur = np.where(cubic[:, None], nxt, cur)
            pts.append(nxt)
        ops["tag"][:, 1 + s] = np.where(cubic, int(OpTag.Cubic), int(OpTag.Quad))
        ops["v"][:, 1 + s, 0:2] = pts[0]
        ops["v"][:, 1 + s, 2:4] = pts[1]
        ops["v"][:, 1 + s, 4:6] = np.where(cubic[:, None], pts[2], _F32(0))
    ops["tag"][:, n_ops - 1] = OpTag.Close
    offsets = np.arange(count + 1, dtype=np.uint64) * np.uint64(n_ops)
    rules = (idx & np.uint64(1)).astype(np.uint8)
    return np.ascontiguousarray(ops.reshape(-1)), offsets, rules


# ---- config 5: one huge raster, many closed polygons --------------------------
def random_polygons(first, count, vertices=64, size=32768, extent=2048):
    """Sub-figures first..first+count-1 of config 5 as ONE op array (Move, Line*(vertices-1), Close each).

    Polygon i: seed stream of (0xB160000 ^ i); centre uniform in [extent/2, size-extent/2)^2; each
    vertex = centre + uniform offset in [-extent/2, extent/2)^2.
    """
    idx = np.arange(first, first + count, dtype=np.uint64)
    seeds = _mix(np.uint64(0xB160000) ^ idx)
    d = _draws(seeds, 2 + 2 * vertices)
    half = _F32(extent / 2.0)
    cspan = _F32(size - extent)
    cx = _unit(d[:, 0]) * cspan + half
    cy = _unit(d[:, 1]) * cspan + half
    ops = np.zeros((count, vertices + 1), dtype=OP_DTYPE)
    for v in range(vertices):
        x = cx + (_unit(d[:, 2 + 2 * v]) * _F32(extent) - half)
        y = cy + (_unit(d[:, 3 + 2 * v]) * _F32(extent) - half)
        ops["tag"][:, v] = OpTag.Move if v == 0 else OpTag.Line
        ops["v"][:, v, 0] = x.astype(np.float32)
        ops["v"][:, v, 1] = y.astype(np.float32)
    ops["tag"][:, vertices] = OpTag.Close
    return np.ascontiguousarray(ops.reshape(-1))


def fill_pixels(width, height, top_row, row_begin=0, row_end=None):
    """Pixel-count convention of SURVEY §8d: the reference resolves every pixel of rows [max(top_row,0), H)."""
    row_end = height if row_end is None else row_end
    first = max(int(top_row), 0, row_begin)
    return width * max(0, row_end - first)
