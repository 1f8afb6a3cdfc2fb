"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Bit-exact everywhere: Fixed vertices, (dir, top_row), Matte8 bytes, and also
Graya8p / Rgba8p pixels (the north_star tolerance for composited pixels is
+-1 LSB; both sides implement the same recalled pix arithmetic, so the tests
assert equality and would report the max deviation if it ever appeared).
"""
import numpy as np
import pytest

import oracle
from footile_b200 import Batch, FillRule, Format, JoinStyle, Path2D, PathOp, Plotter, Raster, debug_accumulate, scenes
from footile_b200.path import OP_DTYPE, OpTag

pytestmark = pytest.mark.gpu

FMT = {Format.Matte8: oracle.MATTE8, Format.Graya8p: oracle.GRAYA8P, Format.Rgba8p: oracle.RGBA8P}


def both(w, h, fmt, init=None, transform=None, tol=None, join=None, **okw):
    ras = Raster(w, h, fmt, init)
    g = Plotter(ras)
    o = oracle.Plotter(w, h, FMT[fmt], init=None if init is None else ras.pixels, **okw)
    if transform is not None:
        g.set_transform(transform)
        o.set_transform(transform)
    if tol is not None:
        g.set_tolerance(tol)
        o.set_tolerance(tol)
    if join is not None:
        g.set_join(join)
        o.set_join(join.kind, join.limit)
    return g, o


def assert_same(g, o, what=""):
    a, b = g.raster().pixels, o.raster()
    if not np.array_equal(a, b):
        d = np.abs(a.astype(int) - b.astype(int))
        ys, xs = np.nonzero(d)
        raise AssertionError("%s: %d bytes differ, max |d|=%d, first at row %d byte %d (gpu %d, oracle %d)" % (
            what, len(ys), d.max(), ys[0], xs[0], a[ys[0], xs[0]], b[ys[0], xs[0]]))


def poly(points, close=True):
    p = Path2D().absolute().move_to(*points[0])
    for q in points[1:]:
        p = p.line_to(*q)
    if close:
        p = p.close()
    return p.finish()


# ---- kernel (d) alone: imgbuf.rs KATs --------------------------------------
def test_accumulate_kats():  # imgbuf.rs:205-232
    b = np.zeros(3000, dtype=np.int16); b[0] = 200
    assert (debug_accumulate(FillRule.NonZero, b) == 200).all()
    d = np.zeros(5000, dtype=np.int16); d[0] = 300
    assert (debug_accumulate(FillRule.NonZero, d) == 255).all()
    e = np.zeros(3000, dtype=np.int16); e[0] = 300
    assert (debug_accumulate(FillRule.EvenOdd, e) == 212).all()


def test_accumulate_random_rows_vs_oracle():
    rng = np.random.default_rng(5)
    for n in (1, 3, 8, 100, 127, 128, 129, 1000, 4096, 5001):
        src = rng.integers(-600, 600, (7, n)).astype(np.int16)
        src[rng.random((7, n)) < 0.7] = 0
        src[0, 0] = 32767  # forces i16 wrap-around in the running sum
        src[0, n // 2] = 32767
        for rule in (FillRule.NonZero, FillRule.EvenOdd):
            got = debug_accumulate(rule, src)
            for r in range(7):
                exp, _ = oracle.accumulate(int(rule), src[r], simd=True)
                assert np.array_equal(got[r], exp), (n, rule, r)


# ---- fig.rs raster KATs through the full device pipeline -------------------
KATS = [  # (w, h, points, expected) — src/fig.rs:723-794
    (9, 1, [(0, 0), (9, 1), (0, 1)], [242, 213, 185, 156, 128, 100, 71, 43, 14]),
    (3, 3, [(-1, 0), (-1, 3), (3, 1.5)], [112, 16, 0, 255, 224, 32, 112, 16, 0]),
    (1, 3, [(0.5, 0), (0.5, 1.5), (1, 3), (1, 0)], [128, 117, 43]),
    (3, 3, [(1.5, 0), (1.5, 1.5), (2, 3), (3, 3), (3, 0)], [0, 128, 255, 0, 117, 255, 0, 43, 255]),
    (9, 1, [(0, 0), (0, 0.3), (9, 0)], [73, 64, 56, 47, 39, 30, 22, 13, 4]),
]


@pytest.mark.parametrize("w,h,pts,exp", KATS)
def test_fig_kats(w, h, pts, exp):
    g = Plotter(Raster(w, h, Format.Matte8))
    g.fill(FillRule.NonZero, poly(pts, close=False), (255,))
    assert g.raster().as_u8_slice().tolist() == exp


def test_fig_3x3_rgba():  # fig.rs:702-721
    g = Plotter(Raster(3, 3, Format.Rgba8p))
    g.fill(FillRule.NonZero, poly([(1, 2), (1, 3), (2, 3), (2, 2)], close=False), (99, 99, 99, 255))
    exp = np.zeros((3, 12), dtype=np.uint8)
    exp[2, 4:8] = (99, 99, 99, 255)
    assert np.array_equal(g.raster().pixels, exp)


def test_plotter_overlapping():  # plotter.rs:389-403 (smoke in the reference; compared with the oracle here)
    path = (Path2D().absolute().move_to(8.0, 4.0).line_to(8.0, 3.0).cubic_to(8.0, 3.0, 8.0, 3.0, 9.0, 3.75)
            .line_to(8.0, 3.75).line_to(8.5, 3.75).line_to(8.5, 3.5).finish())
    g, o = both(16, 16, Format.Matte8)
    g.fill(FillRule.NonZero, path, (255,))
    o.fill(oracle.NONZERO, path, (255,))
    assert_same(g, o)


# ---- (a) flattening: Fixed vertices bit-exact -------------------------------
def random_path(rng, size, n_seg, closed_prob=0.5):
    p = Path2D().absolute()
    p = p.move_to(*rng.uniform(0, size, 2))
    for _ in range(n_seg):
        k = rng.integers(0, 6)
        if k == 0:
            p = p.line_to(*rng.uniform(-0.2 * size, 1.2 * size, 2))
        elif k in (1, 2):
            p = p.quad_to(*rng.uniform(-0.2 * size, 1.2 * size, 4))
        elif k in (3, 4):
            p = p.cubic_to(*rng.uniform(-0.2 * size, 1.2 * size, 6))
        else:
            if rng.random() < closed_prob:
                p = p.close()
            if rng.random() < 0.7:
                p = p.move_to(*rng.uniform(0, size, 2))
    return p.finish()


@pytest.mark.parametrize("seed", range(6))
def test_flatten_vertices_bit_exact(seed):
    rng = np.random.default_rng(100 + seed)
    size = [16, 64, 256, 1024, 4096, 20000][seed]
    tr = [None, [2, 0, 0, 0, 2, 0], [0.7, -0.3, 5.5, 0.2, 1.3, -3.25]][seed % 3]
    tol = [None, 0.01, 1.0][seed % 3]
    g, o = both(32, 32, Format.Matte8, transform=tr, tol=tol, vid_cap=1 << 30)
    for _ in range(8):
        ops = random_path(rng, size, int(rng.integers(1, 40)))
        gx, gs = g.debug_flatten(ops)
        ox, os_ = o.debug_flatten(ops)
        assert np.array_equal(gs, os_)
        assert np.array_equal(gx, ox)


# ---- (b): Edge::new on the device against the reference's (fig.rs:179-210), with the winding sign (fig.rs:286) ----
@pytest.mark.parametrize("seed", range(4))
def test_edges_bit_exact(seed):
    rng = np.random.default_rng(400 + seed)
    for it in range(12):
        size = int(rng.choice([16, 64, 300, 2000]))
        path = random_path(rng, size, int(rng.integers(2, 24)))
        rule = int(rng.integers(0, 2))
        g, o = both(size, size, Format.Matte8)
        g.fill(rule, path, (255,))
        got = g.debug_edges()
        exp = o.debug_edges(path)
        assert got.shape == exp.shape, (it, got.shape, exp.shape)
        key = lambda a: a[np.lexsort(a.T[::-1])]  # the device builds one edge per ring slot, the oracle walks the ring: compare as sets
        assert np.array_equal(key(got), key(exp)), "it %d" % it


def test_flatten_degenerate_sequences():
    # de-dup, closing-point pop, Move/Move, Line right after Close, leading Close
    cases = [
        [PathOp.Move(1, 1), PathOp.Move(2, 2), PathOp.Line(5, 2), PathOp.Line(5, 2), PathOp.Line(2, 9), PathOp.Line(2, 2)],
        [PathOp.Close(), PathOp.Line(3, 3), PathOp.Line(9, 3), PathOp.Line(9, 9), PathOp.Close(), PathOp.Line(4, 1), PathOp.Line(7, 5)],
        [PathOp.Move(1, 1), PathOp.Close(), PathOp.Close(), PathOp.Quad(5, 0, 9, 9), PathOp.PenWidth(3), PathOp.Cubic(1, 9, 0, 5, 0, 0)],
        [PathOp.Line(4, 4)],
        [PathOp.Move(4, 4)],
        [PathOp.PenWidth(2.0)],
        [PathOp.Move(0, 0), PathOp.Cubic(0, 0, 0, 0, 0, 0), PathOp.Line(0, 0), PathOp.Close()],
        [PathOp.Move(2, 2), PathOp.Line(6, 2), PathOp.Line(6, 6), PathOp.Line(2, 2), PathOp.Line(2, 2), PathOp.Move(1, 1), PathOp.Line(1, 7), PathOp.Line(3, 7)],
    ]
    for ops in cases:
        g, o = both(12, 12, Format.Matte8)
        gx, gs = g.debug_flatten(ops)
        ox, os_ = o.debug_flatten(ops)
        assert np.array_equal(gs, os_) and np.array_equal(gx, ox), ops
        g.fill(FillRule.NonZero, ops, (255,))
        o.fill(oracle.NONZERO, ops, (255,))
        assert_same(g, o, str(ops))
        assert g.debug_last_fill() == o.last_info() or o.last_info()["n_points"] == 0


def test_empty_and_noop_paths():
    init = np.full((5, 7), 9, dtype=np.uint8)
    g = Plotter(Raster(7, 5, Format.Matte8, init))
    g.fill(FillRule.NonZero, [], (255,))
    g.fill(FillRule.EvenOdd, [PathOp.Move(3, 3)], (255,))  # single point: popped, nothing drawn (fig.rs:491)
    assert np.array_equal(g.raster().pixels, init)


# ---- (b)+(c)+(d): random figures, all formats, both rules -------------------
@pytest.mark.parametrize("route", ["default", "bins", "tiles64"])
@pytest.mark.parametrize("fmt", [Format.Matte8, Format.Rgba8p, Format.Graya8p])
def test_random_polygons_vs_oracle(fmt, route, monkeypatch):
    # default: the one-launch small fill.  bins: the general pipeline, jobs of more than 8 edges through raster_bins.
    # tiles64: the general pipeline with FTL_DIRECT_MAX=64, which keeps raster_tiles' own scatter of 9..64 edges covered.
    if route != "default":
        monkeypatch.setenv("FTL_NO_SMALL", "1")
        monkeypatch.setenv("FTL_DIRECT_MAX", "8" if route == "bins" else "64")
    _random_polygons_vs_oracle(fmt)


def _random_polygons_vs_oracle(fmt):
    rng = np.random.default_rng(7 + int(fmt))
    for it in range(60):
        w = int(rng.choice([1, 3, 8, 17, 40, 130, 257]))
        h = int(rng.choice([1, 3, 8, 17, 40, 97]))
        bpp = {Format.Matte8: 1, Format.Graya8p: 2, Format.Rgba8p: 4}[fmt]
        init = rng.integers(0, 256, (h, w * bpp)).astype(np.uint8)
        g, o = both(w, h, fmt, init=init)
        for layer in range(2):
            ops = []
            for _ in range(int(rng.integers(1, 4))):
                n = int(rng.integers(1, 9))
                snap = rng.random() < 0.4
                pts = np.stack([rng.uniform(-w, 2 * w, n), rng.uniform(-h / 2, 1.5 * h, n)], axis=1)
                if snap:
                    pts = np.round(pts * 2) / 2
                ops += list(poly([tuple(p) for p in pts], close=rng.random() < 0.7))
            rule = int(rng.integers(0, 2))
            clr = rng.integers(0, 256, 4).astype(np.uint8)
            clr[:3] = np.minimum(clr[:3], clr[3])
            if fmt == Format.Graya8p:
                clr[0] = min(clr[0], clr[1])
            g.fill(rule, ops, clr)
            o.fill(rule, ops, clr)
            assert g.debug_last_fill() == o.last_info()
            assert_same(g, o, "it %d layer %d" % (it, layer))


# Tiles with at most 8 edge slots on rasters whose width is a multiple of 16 take the analytic rows of
# the tile kernel (no shared-memory scatter); rows where two spans share a 16-pixel group, or a span is
# long (shallow edges), fall back to the shared-memory path inside the same tile.
@pytest.mark.parametrize("fmt", [Format.Matte8, Format.Rgba8p, Format.Graya8p])
def test_analytic_rows_vs_oracle(fmt):
    rng = np.random.default_rng(101 + int(fmt))
    bpp = {Format.Matte8: 1, Format.Graya8p: 2, Format.Rgba8p: 4}[fmt]
    for it in range(120):
        w = int(rng.choice([16, 32, 48, 128, 272, 1040]))
        h = int(rng.choice([1, 4, 5, 8, 19, 64]))
        init = rng.integers(0, 256, (h, w * bpp)).astype(np.uint8)
        g, o = both(w, h, fmt, init=init)
        for layer in range(3):
            kind = int(rng.integers(0, 5))
            n = int(rng.integers(3, 9))  # <= 8 vertices in the whole job
            if kind == 0:  # anywhere, partly outside
                pts = np.stack([rng.uniform(-0.5 * w, 1.5 * w, n), rng.uniform(-0.5 * h, 1.5 * h, n)], axis=1)
            elif kind == 1:  # steep edges: short spans
                xs = np.sort(rng.uniform(0, w, n))
                pts = np.stack([xs, rng.uniform(-1, h + 1, n)], axis=1)
            elif kind == 2:  # shallow edges: long spans, fall back
                pts = np.stack([rng.uniform(-w, 2 * w, n), rng.uniform(0, min(h, 3.0), n)], axis=1)
            elif kind == 3:  # snapped to half pixels, coincident edges
                pts = np.round(np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], axis=1) * 2) / 2
            else:  # two small sub-figures close to each other (spans sharing a group)
                c = rng.uniform(0, w)
                pts = np.stack([c + rng.uniform(-12, 12, n), rng.uniform(0, h, n)], axis=1)
            if kind == 4 and n >= 6:
                ops = list(poly([tuple(q) for q in pts[:3]])) + list(poly([tuple(q) for q in pts[3:]]))
            else:
                ops = list(poly([tuple(q) for q in pts], close=rng.random() < 0.8))
            rule = int(rng.integers(0, 2))
            clr = rng.integers(0, 256, 4).astype(np.uint8)
            if rng.random() < 0.4:  # opaque colour: the no-read path
                clr[3] = 255
                clr[1] = 255 if fmt == Format.Graya8p else clr[1]
            clr[:3] = np.minimum(clr[:3], clr[3])
            if fmt == Format.Graya8p:
                clr[0] = min(clr[0], clr[1])
            g.fill(rule, ops, clr)
            o.fill(rule, ops, clr)
            assert g.debug_last_fill() == o.last_info()
            assert_same(g, o, "it %d layer %d kind %d %dx%d" % (it, layer, kind, w, h))


def test_negative_top_row_shift():
    ops = poly([(1.0, -2.5), (6.0, 3.0), (0.5, 5.0)])
    g, o = both(8, 8, Format.Matte8)
    g.fill(FillRule.NonZero, ops, (255,))
    o.fill(oracle.NONZERO, ops, (255,))
    assert g.debug_last_fill()["top_row"] == -3
    assert_same(g, o)


def test_rows_above_top_untouched_and_below_overwritten():
    init = np.full((16, 16), 77, dtype=np.uint8)
    g, o = both(16, 16, Format.Matte8, init=init)
    ops = poly([(4, 5.5), (12, 6), (8, 9)])
    g.fill(FillRule.NonZero, ops, (255,))
    o.fill(oracle.NONZERO, ops, (255,))
    r = g.raster().pixels
    assert (r[:5] == 77).all() and (r[10:] == 0).all()
    assert_same(g, o)


@pytest.mark.parametrize("rule", [FillRule.NonZero, FillRule.EvenOdd])
def test_curved_paths_vs_oracle(rule):
    rng = np.random.default_rng(31 + int(rule))
    for size in (64, 200, 512):
        for _ in range(6):
            ops = random_path(rng, size, int(rng.integers(3, 30)))
            g, o = both(size, size, Format.Matte8)
            g.fill(rule, ops, (255,))
            o.fill(int(rule), ops, (255,))
            assert g.debug_last_fill() == o.last_info()
            assert_same(g, o)


# ---- config 1: fishy --------------------------------------------------------
def test_config1_fishy_example_rgba():  # examples/fishy.rs:9-31
    fish, eye = scenes.fishy_example()
    g, o = both(128, 128, Format.Rgba8p)
    for p in (g, o):
        p.fill(0, fish, (127, 96, 96, 255))
        p.stroke(fish, (255, 208, 208, 255))
        p.stroke(eye, (0, 0, 0, 255))
    assert_same(g, o)
    assert g.raster().pixels.any()
    assert g.pen_width() == o.pen_width() == 2.0  # PenWidth persists (plotter.rs:151-153)


@pytest.mark.parametrize("size", [16, 256])
def test_config1_fishy_bench(size):  # benches/fishyb.rs:10-39 (scale 2; stroke exercises the double transform)
    path = scenes.fishy_bench()
    g, o = both(size, size, Format.Matte8, transform=[2, 0, 0, 0, 2, 0])
    g.fill(FillRule.NonZero, path, (255,))
    o.fill(oracle.NONZERO, path, (255,))
    assert_same(g, o, "fill")
    g, o = both(size, size, Format.Matte8, transform=[2, 0, 0, 0, 2, 0])
    g.stroke(path, (255,))
    o.stroke(path, (255,))
    assert_same(g, o, "stroke")


# ---- config 2: heptagram -----------------------------------------------------
@pytest.mark.parametrize("rule", [FillRule.NonZero, FillRule.EvenOdd])
@pytest.mark.parametrize("size", [100, 4096])
def test_config2_heptagram(rule, size):
    path = scenes.heptagram_abs()
    tr = scenes.heptagram_transform(size)
    g, o = both(size, size, Format.Matte8, transform=tr)
    g.fill(rule, path, (255,))
    o.fill(int(rule), path, (255,))
    assert g.debug_last_fill() == o.last_info()
    assert_same(g, o)
    r = g.raster().pixels
    assert r.max() == 255 and (r[: g.debug_last_fill()["top_row"]] == 0).all()


def test_config2_heptagram_example_relative():  # examples/heptagram.rs:12-24 (relative builder, EvenOdd, 100x100)
    path = scenes.heptagram()
    g, o = both(100, 100, Format.Matte8, transform=[50, 0, 25, 0, 50, 25])
    g.fill(FillRule.EvenOdd, path, (255,))
    o.fill(oracle.EVENODD, path, (255,))
    assert_same(g, o)


# ---- config 3: strokes --------------------------------------------------------
@pytest.mark.parametrize("join", [JoinStyle.Miter(4.0), JoinStyle.Bevel, JoinStyle.Round])
def test_config3_stroke_scenes_small(join):  # examples/*.rs at their own size
    for name, path in scenes.stroke_scenes(1.0).items():
        g, o = both(128, 128, Format.Matte8, join=join)
        go = g.debug_stroke_ops(path)
        oo = o.debug_stroke_ops(path)
        assert len(go) == len(oo) and go.tobytes() == oo.tobytes(), name
        g.stroke(path, (255,))
        o.stroke(path, (255,))
        assert_same(g, o, name)
        assert g.raster().pixels.any(), name


@pytest.mark.parametrize("join", [JoinStyle.Round, JoinStyle.Miter(4.0), JoinStyle.Bevel])
def test_config3_strokes_4k_rgba(join):  # 3840x2160 Rgba8p over (64,128,64,255), scenes x30
    w, h = 3840, 2160
    base = Raster.with_color(w, h, Format.Rgba8p, (64, 128, 64, 255)).pixels
    g, o = both(w, h, Format.Rgba8p, init=base, join=join)
    for name, path in scenes.stroke_scenes(30.0).items():
        g.stroke(path, (255, 255, 0, 255))
        o.stroke(path, (255, 255, 0, 255))
    assert_same(g, o)


def _round_rs(join):
    """examples/round.rs:9-17 with the given join: 100x100 Matte8, pen width 40, (10,60) -> (60,60) -> (60,10)."""
    path = scenes.stroke_scenes(1.0)["round"]
    g, o = both(100, 100, Format.Matte8, join=join)
    g.stroke(path, (255,))
    o.stroke(path, (255,))
    assert_same(g, o, str(join))
    return g.raster().pixels


def test_stroke_join_semantics_outer_corner():
    """A guard on the recalled pointy semantics (right(), angle_rel, intersection: SURVEY App. B): with the
    wrong sign of right() GPU and oracle would still agree, but the join would land on the INNER corner.
    The path turns at (60,60); the outer corner is towards (80,80), the inner one towards (40,40)
    (stroker.rs:301-309 round_point, 380-396 miter / bevel)."""
    rnd = _round_rs(JoinStyle.Round)
    # round join: a disc of radius 20 about (60,60) - (73,73) is 18.4 away (inside), (78,78) 25.5 away (outside)
    assert rnd[73, 73] == 255 and rnd[78, 78] == 0
    assert rnd[66, 76] == 255 and rnd[76, 66] == 255  # inside the disc (17.7 from the corner), outside the bevel chord x + y = 140
    mit = _round_rs(JoinStyle.Miter(4.0))
    assert mit[78, 78] == 255 and mit[73, 73] == 255   # the miter tip reaches (80,80)
    bev = _round_rs(JoinStyle.Bevel)
    assert bev[73, 73] == 0 and bev[68, 68] == 255     # the bevel is the chord (60,80)-(80,60): x + y = 140
    for img in (rnd, mit, bev):
        assert img[45, 45] == 255 and img[35, 35] == 0  # inner corner: inside both strokes at (45,45), nothing at (35,35)
        assert img[70, 20] == 255 and img[20, 70] == 255 and img[5, 5] == 0  # the two straight runs (rows = y, columns = x)
        assert img[60, 85] == 0 and img[85, 60] == 0    # butt ends are not extended past the outer edges (x = 80, y = 80)


# ---- config 4: batch of random curve paths -------------------------------------
def fnv_expected(img):
    """Host replica of ftl_batch_checksums: 256 interleaved FNV-1a lanes folded in order."""
    b = np.ascontiguousarray(img).ravel()
    assert b.size % 256 == 0
    lanes = b.reshape(-1, 256).astype(np.uint64)
    h = np.full(256, 0xcbf29ce484222325, dtype=np.uint64)
    prime = np.uint64(0x100000001b3)
    with np.errstate(over="ignore"):
        for row in lanes:
            h = (h ^ row) * prime
        g = np.uint64(0xcbf29ce484222325)
        for v in h:
            for k in range(8):
                g = (g ^ ((v >> np.uint64(8 * k)) & np.uint64(0xFF))) * prime
    return int(g)


def test_config4_batch_matches_oracle():
    n, size = 48, 512
    ops, offs, rules = scenes.random_curve_paths(1000, n)
    b = Batch(size, size, Format.Matte8, n)
    b.fill(ops, offs, rules=rules)
    got = b.read()
    sums = b.checksums()
    for j in range(n):
        o = oracle.Plotter(size, size, oracle.MATTE8)
        o.fill(int(rules[j]), ops[int(offs[j]): int(offs[j + 1])], (255,))
        exp = o.raster()
        assert np.array_equal(got[j], exp), "path %d" % j
        if j % 8 == 0:
            assert int(sums[j]) == fnv_expected(exp)
    assert got.any()


def test_config4_full_size_100k_paths_checksums():
    """BASELINE config 4 at its full size: all 100 000 random 64-curve paths, each into its own 512^2 Matte8
    raster (25.6 GiB resident in HBM), every raster's device checksum against the oracle's."""
    import os
    total, size, chunk = 100_000, 512, 12_500
    threads = min(32, os.cpu_count() or 1)
    b = Batch(size, size, Format.Matte8, chunk)
    for first in range(0, total, chunk):
        ops, offs, rules = scenes.random_curve_paths(first, chunk)
        b.clear()  # rows above a figure's top row keep their content (fig.rs:497): every chunk starts from zero
        b.fill(ops, offs, rules=rules)
        got = b.checksums()
        exp = oracle.batch_fill_checksums(size, size, oracle.MATTE8, ops, offs, rules=rules, clr=(255,), threads=threads)
        bad = np.nonzero(got != exp)[0]
        assert bad.size == 0, "paths %s differ" % (first + bad[:8])


def test_config4_batch_transforms_colors_rgba():
    n, size = 12, 64
    rng = np.random.default_rng(3)
    paths = [random_path(rng, size, 12) for _ in range(n)]
    ops, offs = Batch.pack(paths)
    rules = rng.integers(0, 2, n).astype(np.uint8)
    tr = np.tile(np.array([1, 0, 0, 0, 1, 0], dtype=np.float32), (n, 1))
    tr[:, 2] = rng.uniform(-5, 5, n)
    tr[:, 0] = rng.uniform(0.5, 1.5, n)
    colors = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    colors[:, :3] = np.minimum(colors[:, :3], colors[:, 3:4])
    b = Batch(size, size, Format.Rgba8p, n)
    b.fill(ops, offs, rules=rules, transforms=tr, colors=colors)
    got = b.read()
    for j in range(n):
        o = oracle.Plotter(size, size, oracle.RGBA8P)
        o.set_transform(tr[j])
        o.fill(int(rules[j]), paths[j], colors[j])
        assert np.array_equal(got[j], o.raster()), j


def test_batch_replay_is_idempotent_for_matte():
    ops, offs, rules = scenes.random_curve_paths(5, 4, segments=8, size=128)
    b = Batch(128, 128, Format.Matte8, 4)
    b.upload(ops, offs, rules=rules)
    b.run()
    first = b.read().copy()
    b.run().run()
    assert np.array_equal(first, b.read())


# ---- config 5: many sub-figures in one fill, row bands ---------------------------
def test_config5_many_subfigures_and_bands():
    size, n_poly = 2048, 600
    ops = scenes.random_polygons(0, n_poly, vertices=64, size=size, extent=256)
    for rule in (FillRule.NonZero, FillRule.EvenOdd):
        o = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True)
        o.fill(int(rule), ops, (255,))
        exp = o.raster()
        g = Plotter(Raster(size, size, Format.Matte8))
        g.fill(rule, ops, (255,))
        assert g.debug_last_fill() == o.last_info()
        assert np.array_equal(g.raster().pixels, exp)
        # the same fill split into 4 row bands, each on its own handle (the multi-GPU layout)
        for k in range(4):
            r0, r1 = k * size // 4, (k + 1) * size // 4
            gb = Plotter(Raster(size, size, Format.Matte8), rows=(r0, r1))
            gb.fill(rule, ops, (255,))
            assert np.array_equal(gb.raster().pixels, exp[r0:r1]), (rule, k)


_C5 = {}


def _config5_full(rule):
    """BASELINE.json configs[4] at FULL size: one 32768x32768 Matte8 raster, 160 000 closed 64-gons (10.24 M
    edges) in ONE fill (fig.rs:480-502).  Cached per rule: (ops, device raster as a host array, fill info)."""
    if rule not in _C5:
        size = 32768
        if "ops" not in _C5:
            _C5["ops"] = scenes.random_polygons(0, 160000, vertices=64, size=size, extent=2048)
        g = Plotter.with_clear(size, size, Format.Matte8)
        g.fill(rule, _C5["ops"], (255,))
        _C5[rule] = (g.raster().pixels, g.debug_last_fill())
        del g
    return _C5["ops"], _C5[rule][0], _C5[rule][1]


@pytest.mark.parametrize("rule", [FillRule.EvenOdd, FillRule.NonZero])
def test_config5_full_size_stripes_vs_oracle(rule):
    """Seeded 16-row stripes of the full-size fill against the order-free oracle (u32 vertex ids): every stripe
    walks all 10.24 M edges on the CPU, so 8 stripes per rule is what a few seconds allow."""
    size, rows = 32768, 16
    ops, full, info = _config5_full(rule)
    rng = np.random.default_rng(5 + int(rule))
    starts = [0, size - rows] + [int(r) for r in rng.integers(0, size - rows, 6)]
    o = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True)
    for r0 in starts:
        o.set_rows(r0, r0 + rows)
        o.fill(int(rule), ops, (255,))
        assert o.last_info() == info
        exp = o.raster()[r0: r0 + rows]
        got = full[r0: r0 + rows]
        assert np.array_equal(exp, got), "rule %s rows %d..%d: %d bytes differ" % (rule, r0, r0 + rows, int((exp != got).sum()))
    # size-independent property of the whole raster: hundreds of overlapping self-intersecting polygons make the
    # winding number a symmetric random walk.  NonZero: negative sums clamp to 0 (imgbuf.rs:54-66), so about half of
    # the pixels are covered; EvenOdd: ~6.5 edge crossings per pixel leave almost no pixel with zero coverage.
    frac = np.count_nonzero(full[::64]) / full[::64].size
    assert (0.4 < frac < 0.6) if rule == FillRule.NonZero else (0.9 < frac < 1.0), frac


def test_config5_full_size_row_bands_equal_unsplit():
    """The multi-GPU layout of config 5: 4 row bands on their own handles reproduce the unsplit raster."""
    size = 32768
    ops, full, info = _config5_full(FillRule.EvenOdd)
    for k in range(4):
        r0, r1 = k * size // 4, (k + 1) * size // 4
        gb = Plotter.with_clear(size, size, Format.Matte8, rows=(r0, r1))
        gb.fill(FillRule.EvenOdd, ops, (255,))
        binfo = gb.debug_last_fill()  # a band flattens only the sub-figures it needs: n_points differs, (dir, top_row) must not
        assert (binfo["dir"], binfo["top_row"]) == (info["dir"], info["top_row"])
        assert np.array_equal(gb.raster().pixels, full[r0:r1]), k
        del gb
    _C5.clear()


def test_sequential_oracle_agrees_on_many_subfigures():
    size = 512
    ops = scenes.random_polygons(7, 40, vertices=16, size=size, extent=128)
    a = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True).fill(0, ops, (255,)).raster()
    b = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=False).fill(0, ops, (255,)).raster()
    assert np.array_equal(a, b)


# ---- wide raster (single row tile per CTA) ---------------------------------------
def test_wide_raster_32768():
    w, h = 32768, 24
    ops = poly([(100.5, 1.25), (32000.0, 3.0), (16000.0, 22.5), (5.0, 20.0)])
    g, o = both(w, h, Format.Matte8)
    g.fill(FillRule.NonZero, ops, (255,))
    o.fill(oracle.NONZERO, ops, (255,))
    assert_same(g, o)


def test_error_paths():
    from footile_b200 import FootileError
    g = Plotter(Raster(8, 8, Format.Matte8))
    with pytest.raises(FootileError):
        g.fill(FillRule.NonZero, [PathOp.Move(float("nan"), 0), PathOp.Line(1, 1)], (255,))
    with pytest.raises(FootileError):
        g.set_transform([float("inf"), 0, 0, 0, 1, 0])
    with pytest.raises(FootileError):
        g.fill(7, [PathOp.Move(0, 0)], (255,))
    g.set_tolerance(-5.0)  # clamps to 0.01 like the reference (plotter.rs:133-137)
    g.fill(FillRule.NonZero, poly([(1, 1), (6, 1), (3, 6)]), (255,))
    assert g.raster().pixels.any()


# ---- speculative (sync-free) replay: scratch overflow is detected and the call repeated -------------
def test_speculative_capacity_overflow_is_recovered():
    rng = np.random.default_rng(11)
    g, o = both(256, 256, Format.Rgba8p, init=np.full((256, 1024), 40, dtype=np.uint8))
    small = poly([(10, 10), (50, 12), (30, 60)])
    big = random_path(rng, 256, 60)  # needs far more vertices / bin entries than `small` sized the buffers for
    for ops, clr in ((small, (200, 10, 10, 255)), (big, (10, 90, 10, 128)), (small, (0, 0, 50, 60)), (big, (9, 9, 9, 9))):
        g.fill(FillRule.NonZero, ops, clr)  # non-idempotent blend: a repeated or dropped pass would show
        o.fill(oracle.NONZERO, ops, clr)
    assert_same(g, o)


def test_clear_after_an_overflowed_replay_stays_clear():
    """a speculative replay that overflowed draws nothing and is repeated later; a clear issued in
    between must come AFTER that repeat.  small fill (sizes the buffers) -> big fill (overflows) -> clear -> read."""
    rng = np.random.default_rng(3)
    small = poly([(10, 10), (50, 12), (30, 60)])
    big = random_path(rng, 256, 60)
    b = Batch(256, 256, Format.Matte8, 1)
    one = np.array([0], dtype=np.uint64)
    b.fill(small, np.array([0, len(small)], dtype=np.uint64))
    b.fill(big, np.array([0, len(big)], dtype=np.uint64))
    b.clear()
    assert not b.read().any()
    # and a loop of clear(); run() on an overflowing job set leaves exactly one fill in the raster
    b2 = Batch(256, 256, Format.Rgba8p, 1)
    clr = np.array([[10, 90, 10, 128]], dtype=np.uint8)
    b2.fill(small, np.array([0, len(small)], dtype=np.uint64), colors=clr)
    b2.upload(big, np.array([0, len(big)], dtype=np.uint64), colors=clr)
    for _ in range(3):
        b2.clear()
        b2.run()
    o = oracle.Plotter(256, 256, oracle.RGBA8P)
    o.fill(oracle.NONZERO, big, (10, 90, 10, 128))
    assert np.array_equal(b2.read()[0], o.raster())
    del one


def test_many_async_replays_then_read():
    ops, offs, rules = scenes.random_curve_paths(77, 6, segments=16, size=128)
    b = Batch(128, 128, Format.Matte8, 6)
    b.upload(ops, offs, rules=rules)
    for _ in range(150):  # more than the pending-check ring holds
        b.run()
    got = b.read()
    for j in range(6):
        o = oracle.Plotter(128, 128, oracle.MATTE8)
        o.fill(int(rules[j]), ops[int(offs[j]): int(offs[j + 1])], (255,))
        assert np.array_equal(got[j], o.raster())


# ---- two GPUs: row bands on two devices gathered over NCCL (skipped on a 1-GPU box) ----------------
def _band_worker(rank, world, port, size, out_dir):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from footile_b200 import sharding
    ops = scenes.random_polygons(0, 300, vertices=32, size=size, extent=300)
    r0, r1 = sharding.band_rows(size, rank, world)
    g = Plotter(Raster(size, size, Format.Matte8), device=rank, rows=(r0, r1))
    g.fill(FillRule.EvenOdd, ops, (255,)).sync()
    ptr, nbytes = g.device_ptr()
    band = sharding.device_tensor(ptr, nbytes, rank).view(r1 - r0, size)
    full = sharding.gather_bands(band, size, size)
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full.cpu().numpy())
    dist.destroy_process_group()


def test_two_gpu_row_bands_nccl_gather(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    size = 1024
    mp.spawn(_band_worker, args=(2, 29611, size, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "full.npy")
    ops = scenes.random_polygons(0, 300, vertices=32, size=size, extent=300)
    exp = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True).fill(1, ops, (255,)).raster()
    assert np.array_equal(got, exp)


# ---- packed read-back (PCIe transport encoding) returns exactly the device bytes ---------------------
def test_packed_readback_roundtrip():
    rng = np.random.default_rng(2)
    w, h = 4096, 2048  # 8 MiB: above the packing threshold
    for kind in ("noise", "mixed", "constant"):
        if kind == "noise":
            img = rng.integers(0, 256, (h, w)).astype(np.uint8)       # incompressible: falls back to the plain copy
        elif kind == "constant":
            img = np.full((h, w), 200, dtype=np.uint8)
        else:
            img = np.zeros((h, w), dtype=np.uint8)
            img[100:900] = 255
            img[:, 1000:1033] = rng.integers(0, 256, (h, 33)).astype(np.uint8)  # literal blocks straddling block borders
            img[5, :] = np.arange(w) % 251
            img[-1, -1] = 7
        g = Plotter(Raster(w, h, Format.Matte8, img))
        assert np.array_equal(g.raster().pixels, img), kind


def test_packed_readback_large():
    """Five pieces of 256 MiB: the packed read-back pipelines the pieces (device classify / PCIe / host expansion)."""
    rng = np.random.default_rng(5)
    w, h = 32768, 40960  # 1.25 GiB = 5 pieces
    img = np.zeros((h, w), dtype=np.uint8)
    img[1000:9000] = 255
    img[30000:30100] = rng.integers(0, 256, (100, w)).astype(np.uint8)
    img[:, 5000:5040] = rng.integers(0, 256, (h, 40)).astype(np.uint8)
    img[-1, -1] = 9
    g = Plotter(Raster(w, h, Format.Matte8, img))
    out = g.raster().pixels
    assert out.shape == img.shape
    for r0 in range(0, h, 4096):  # compare in slices: a full-size temporary is not needed
        assert np.array_equal(out[r0:r0 + 4096], img[r0:r0 + 4096]), r0


# ---- wide + dense: edges are binned per (band, row window); long shallow edges cross several windows ----
@pytest.mark.parametrize("fmt", [Format.Matte8, Format.Rgba8p])
def test_wide_dense_window_bins(fmt):
    rng = np.random.default_rng(99)
    w, h = 9000, 40   # 5 windows of 2048 cells
    ops = []
    for _ in range(60):  # 60 polygons x 5 vertices = 300 edges > DIRECT_MAX, many nearly horizontal
        n = 5
        xs = rng.uniform(-1500, w + 1500, n)
        ys = rng.uniform(-5, h + 5, n)
        if rng.random() < 0.3:
            ys = np.round(ys)
        ops += list(poly([(float(x), float(y)) for x, y in zip(xs, ys)]))
    for rule in (FillRule.NonZero, FillRule.EvenOdd):
        g, o = both(w, h, fmt, vid_cap=1 << 30)
        clr = (90, 60, 30, 200)
        g.fill(rule, ops, clr)
        o.fill(int(rule), ops, clr)
        assert g.debug_last_fill() == o.last_info()
        assert_same(g, o, "rule %d" % rule)


def test_flatten_extreme_curves_hit_the_same_depth_cap():
    # huge control polygons at the minimum tolerance: deep subdivision, same vertex list on both sides
    g, o = both(64, 64, Format.Matte8, tol=0.01, vid_cap=1 << 30)
    ops = [PathOp.Move(-30000, -30000), PathOp.Cubic(32000, -32000, 31000, 32000, -30000, 30000), PathOp.Quad(1e7, -1e7, 5, 5), PathOp.Close()]
    gx, gs = g.debug_flatten(ops)
    ox, os_ = o.debug_flatten(ops)
    assert len(gx) > 1000
    assert np.array_equal(gs, os_) and np.array_equal(gx, ox)


# ---- layered scene: many fills/strokes onto one raster in one pass ---------------------
@pytest.mark.parametrize("fmt", [Format.Rgba8p, Format.Matte8, Format.Graya8p])
def test_fill_layers_equals_sequential_calls(fmt):
    rng = np.random.default_rng(21 + int(fmt))
    w, h = 300, 200
    bpp = {Format.Matte8: 1, Format.Graya8p: 2, Format.Rgba8p: 4}[fmt]
    init = rng.integers(0, 256, (h, w * bpp)).astype(np.uint8)
    g, o = both(w, h, fmt, init=init, join=JoinStyle.Round)
    layers = []
    for k in range(14):
        path = random_path(rng, 250, int(rng.integers(2, 12)))
        clr = rng.integers(0, 256, 4).astype(np.uint8)
        clr[:3] = np.minimum(clr[:3], clr[3])
        if fmt == Format.Graya8p:
            clr[0] = min(clr[0], clr[1])
        if k % 3 == 2:  # a stroke layer: its outline, filled NonZero
            path = np.concatenate([np.array([PathOp.PenWidth(float(rng.uniform(1, 9)))], dtype=OP_DTYPE), path])
            layers.append((FillRule.NonZero, g.stroke_outline(path), clr))
            o.stroke(path, clr)
        else:
            rule = int(rng.integers(0, 2))
            layers.append((rule, path, clr))
            o.fill(rule, path, clr)
    g.fill_layers(layers)
    assert_same(g, o)
    assert g.pen_width() == o.pen_width()


def test_fishy_example_as_one_layered_call():  # examples/fishy.rs:29-31 in one pass
    fish, eye = scenes.fishy_example()
    g, o = both(128, 128, Format.Rgba8p)
    g.fill_layers([(FillRule.NonZero, fish, (127, 96, 96, 255)), (FillRule.NonZero, g.stroke_outline(fish), (255, 208, 208, 255)),
                   (FillRule.NonZero, g.stroke_outline(eye), (0, 0, 0, 255))])
    o.fill(0, fish, (127, 96, 96, 255))
    o.stroke(fish, (255, 208, 208, 255))
    o.stroke(eye, (0, 0, 0, 255))
    assert_same(g, o)


# ---- size-independent properties at BASELINE.json's full sizes ------------------------------------------
def test_property_integer_translation_shifts_the_matte():
    # a figure moved by whole pixels yields the same matte moved by whole pixels; coordinates are multiples of
    # 1/8 px so that both the figure and its translate are exact in f32 and in Fixed
    size, dx, dy = 4096, 37, 101
    rng = np.random.default_rng(4)
    pts = np.round(rng.uniform(300, 3600, (11, 2)) * 8) / 8
    a = Plotter.with_clear(size, size, Format.Matte8).fill(FillRule.EvenOdd, poly([tuple(p) for p in pts]), (255,)).raster().pixels
    b = Plotter.with_clear(size, size, Format.Matte8).fill(FillRule.EvenOdd, poly([(p[0] + dx, p[1] + dy) for p in pts]), (255,)).raster().pixels
    assert a.any()
    top = int(pts[:, 1].min())
    assert np.array_equal(a[top: size - dy, : size - dx], b[top + dy:, dx:])


def test_property_matte_fill_is_idempotent_and_deterministic_full_batch():
    # config 4 at one step of its full size: 4096 paths; refilling changes nothing; two batches agree; a seeded
    # sample of rasters equals the oracle and the device checksums equal the host checksums of those rasters
    n = 4096
    ops, offs, rules = scenes.random_curve_paths(0, n)
    b1 = Batch(512, 512, Format.Matte8, n)
    b1.fill(ops, offs, rules=rules)
    s1 = b1.checksums()
    b1.fill(ops, offs, rules=rules)
    assert np.array_equal(s1, b1.checksums())
    b2 = Batch(512, 512, Format.Matte8, n)
    b2.fill(ops, offs, rules=rules)
    assert np.array_equal(s1, b2.checksums())
    assert len(np.unique(s1)) > n // 2
    for j in np.random.default_rng(0).integers(0, n, 12):
        o = oracle.Plotter(512, 512, oracle.MATTE8)
        o.fill(int(rules[j]), ops[int(offs[j]): int(offs[j + 1])], (255,))
        exp = o.raster()
        assert np.array_equal(b1.read(int(j), 1)[0], exp)
        assert int(s1[j]) == fnv_expected(exp)


# ---- the one-launch path for small fills (small_kernel.cuh) against the general pipeline and the oracle ----------
def _general_path(on):
    import os
    if on:
        os.environ["FTL_NO_SMALL"] = "1"
    else:
        os.environ.pop("FTL_NO_SMALL", None)


@pytest.mark.parametrize("fmt", [Format.Matte8, Format.Rgba8p, Format.Graya8p])
def test_small_fill_equals_general_pipeline_and_oracle(fmt):
    rng = np.random.default_rng(2024)
    for it in range(24):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        if it % 6 == 0:
            w, h = 1024, 37  # four windows per band
        ops = random_path(rng, max(w, h), int(rng.integers(1, 12)))
        rule = FillRule.EvenOdd if it & 1 else FillRule.NonZero
        clr = tuple(int(v) for v in rng.integers(0, 256, 4))
        init = rng.integers(0, 256, (h, w * {Format.Matte8: 1, Format.Graya8p: 2, Format.Rgba8p: 4}[fmt]), dtype=np.uint8)
        imgs, infos, edges = [], [], []
        for general in (False, True):
            _general_path(general)
            try:
                g, o = both(w, h, fmt, init=init, tol=0.3 if it % 3 else 0.05)
                g.fill(rule, ops, clr)
                g.fill(rule, ops, clr)  # twice: blends are not idempotent, and the tile must be left clean
                imgs.append(g.raster().pixels)
                infos.append(g.debug_last_fill())
                edges.append(np.sort(g.debug_edges().view([("f%d" % k, "<i4") for k in range(6)]), axis=0))
            finally:
                _general_path(False)
        o.fill(int(rule), ops, clr)
        o.fill(int(rule), ops, clr)
        assert infos[0] == infos[1] == o.last_info(), it
        assert np.array_equal(edges[0], edges[1]), it
        assert np.array_equal(imgs[0], imgs[1]), it
        assert np.array_equal(imgs[0], o.raster()), it


def test_small_fill_overflow_falls_back_in_order():
    """A fill that does not fit the one-launch kernel (a curve with more than 64 points) draws nothing there, poisons the
    small fills issued after it, and all of them are repeated IN ORDER by the general pipeline at the next blocking call."""
    big = Path2D().absolute().move_to(5, 5).cubic_to(900, 20, 20, 900, 600, 600).close().finish()
    small = poly([(10, 10), (200, 30), (120, 220)])
    g, o = both(640, 640, Format.Rgba8p, tol=0.01)
    seq = [(small, (200, 10, 10, 128)), (big, (10, 200, 10, 128)), (small, (10, 10, 200, 128)), (big, (90, 90, 0, 200)), (small, (0, 50, 50, 77))]
    for ops, clr in seq:
        g.fill(FillRule.NonZero, ops, clr)
        o.fill(oracle.NONZERO, ops, clr)
    assert_same(g, o)
    assert g.debug_last_fill() == o.last_info()


# ---- strokes: host-side flatten == device flatten; batched strokes == one Plotter per stroke ------------------
def test_stroke_flatten_host_equals_device_kernel():
    import os
    rng = np.random.default_rng(99)
    for it in range(12):
        path = random_path(rng, 300, int(rng.integers(2, 14)), closed_prob=0.5)
        width = PathOp.PenWidth(float(rng.uniform(0.5, 12.0)))
        ops = np.concatenate([np.array([width], dtype=OP_DTYPE), path, np.array([PathOp.PenWidth(3.0)], dtype=OP_DTYPE), path[: len(path) // 2]])
        join = [JoinStyle.Miter(4.0), JoinStyle.Bevel, JoinStyle.Round][it % 3]
        g, o = both(320, 320, Format.Matte8, join=join, tol=0.3 if it & 1 else 0.05)
        host = g.debug_stroke_ops(ops)
        os.environ["FTL_DEVICE_STROKE_FLATTEN"] = "1"
        try:
            dev = g.debug_stroke_ops(ops)
        finally:
            os.environ.pop("FTL_DEVICE_STROKE_FLATTEN", None)
        ref = o.debug_stroke_ops(ops)
        assert len(host) == len(dev) == len(ref) and host.tobytes() == dev.tobytes() == ref.tobytes(), it


@pytest.mark.parametrize("join", [JoinStyle.Round, JoinStyle.Miter(4.0)])
def test_batch_stroke_equals_one_plotter_per_stroke(join):
    paths = list(scenes.stroke_scenes(2.0).values()) + [scenes.fishy_bench(), scenes.fishy_example()[0]]
    n = len(paths)
    ops, offs = Batch.pack(paths)
    tr = np.tile(np.array([1, 0, 0, 0, 1, 0], dtype=np.float32), (n, 1))
    tr[-2] = [1.5, 0, 3, 0, 1.5, 2]  # the double transform of Plotter::stroke shows with a non-identity transform
    colors = np.tile(np.array([200, 120, 40, 255], dtype=np.uint8), (n, 1))
    b = Batch(256, 256, Format.Rgba8p, n).set_join(join)
    b.stroke(ops, offs, transforms=tr, colors=colors)
    got = b.read()
    for j, path in enumerate(paths):
        o = oracle.Plotter(256, 256, oracle.RGBA8P)
        o.set_join(join.kind, join.limit)
        o.set_transform(tr[j])
        o.stroke(path, (200, 120, 40, 255))
        assert np.array_equal(got[j], o.raster()), j


# ---- the stroker on the device (stroke_kernels.cuh) against the host stroker and the oracle ------------------
JOINS = [JoinStyle.Miter(4.0), JoinStyle.Bevel, JoinStyle.Round, JoinStyle.Miter(1.5)]


def test_libm_restatement_on_this_box():
    """The same pin as tests/test_host.py::test_libm_restatement_matches_glibc, on the GPU box's own CPU and glibc."""
    from footile_b200.plotter import debug_libm_selftest
    h, a, s_wrong, _ = debug_libm_selftest(1_000_000, 99)
    assert (h, a, s_wrong) == (0, 0, 0)


def test_device_stroker_outline_equals_host_on_the_example_scenes():
    """stroker.rs:204-416 on the device: the outline ops equal the host stroker's (and the oracle's) bit for bit."""
    declined = 0
    for scale in (1.0, 30.0):
        for join in JOINS:
            for name, path in list(scenes.stroke_scenes(scale).items()) + [("fishy", scenes.fishy_example()[0]), ("eye", scenes.fishy_example()[1])]:
                g, o = both(128, 128, Format.Matte8, join=join)
                dev = g.debug_stroke_ops_device(path)
                ref = o.debug_stroke_ops(path)
                if dev is None:
                    declined += 1
                    continue
                assert len(dev) == len(ref) and dev.tobytes() == ref.tobytes(), (name, scale, str(join))
    assert declined == 0, declined  # none of the reference's scenes has a join on a sin threshold


def test_device_stroker_outline_equals_host_on_random_paths():
    rng = np.random.default_rng(2024)
    declined = 0
    n_cases = 60
    for it in range(n_cases):
        parts = []
        for _ in range(int(rng.integers(1, 5))):  # several sub-strokes: Move / Close / PenWidth between them
            parts.append(np.array([PathOp.PenWidth(float(rng.uniform(0.3, 25.0)))], dtype=OP_DTYPE))
            parts.append(random_path(rng, 300, int(rng.integers(1, 14)), closed_prob=0.5))
            if rng.random() < 0.3:
                parts.append(np.array([PathOp.Close()], dtype=OP_DTYPE))  # a second close on the same sub-stroke
        ops = np.concatenate(parts)
        join = JOINS[it % 4]
        tr = None if it % 3 else [1.25, 0.1, 5.0, -0.2, 0.9, 11.0]
        g, o = both(320, 320, Format.Matte8, join=join, tol=0.3 if it & 1 else 0.05, transform=tr)
        dev = g.debug_stroke_ops_device(ops)
        host = g.debug_stroke_ops(ops)
        ref = o.debug_stroke_ops(ops)
        assert host.tobytes() == ref.tobytes(), it
        if dev is None:
            declined += 1
            continue
        assert len(dev) == len(ref) and dev.tobytes() == ref.tobytes(), it
    assert declined <= n_cases // 10, declined


def test_device_stroker_degenerate_inputs():
    P = PathOp
    cases = [
        [P.Move(10, 10)],                                              # one point: no segment, nothing drawn
        [P.Move(10, 10), P.Close()],                                   # joined sub-stroke of one point
        [P.Move(10, 10), P.Line(10, 10), P.Line(40, 10)],              # coincident points are dropped (stroker.rs:209)
        [P.Line(30, 30), P.Line(60, 30), P.Line(60, 60), P.Close(), P.Move(5, 5), P.Line(9, 50)],  # Move after Close un-joins (stroker.rs:230-236)
        [P.Close(), P.Close(), P.Line(3, 3), P.Line(50, 20), P.Close(), P.Close()],
        [P.Move(20, 20), P.Line(80, 20), P.Line(20, 20), P.Close()],   # a reversal and a closing segment of zero length
        [P.Move(20, 20), P.Line(50, 20), P.Line(80, 20), P.Line(80, 50), P.Line(80, 80)],  # collinear points, axis-aligned
        [P.PenWidth(0.0), P.Move(20, 20), P.Line(80, 30), P.Line(30, 80)],  # zero width
        [P.PenWidth(7.0)],
    ]
    for join in JOINS:
        for k, ops in enumerate(cases):
            ops = np.array(ops, dtype=OP_DTYPE)
            g, o = both(100, 100, Format.Matte8, join=join)
            dev = g.debug_stroke_ops_device(ops)
            ref = o.debug_stroke_ops(ops)
            assert dev is not None, (k, str(join))
            assert len(dev) == len(ref) and dev.tobytes() == ref.tobytes(), (k, str(join))


@pytest.mark.parametrize("join", [JoinStyle.Round, JoinStyle.Miter(4.0), JoinStyle.Bevel])
def test_device_stroker_pixels_4k(join, monkeypatch):
    """Plotter::stroke with the stroker forced onto the device: flatten -> outline -> fill without leaving HBM."""
    monkeypatch.setenv("FTL_DEVICE_STROKE", "1")
    w, h = 3840, 2160
    base = Raster.with_color(w, h, Format.Rgba8p, (64, 128, 64, 255)).pixels
    g, o = both(w, h, Format.Rgba8p, init=base, join=join)
    from footile_b200 import launch_count
    before = launch_count()
    for name, path in scenes.stroke_scenes(30.0).items():
        g.stroke(path, (255, 255, 0, 255))
        o.stroke(path, (255, 255, 0, 255))
    assert launch_count() - before > 30 * len(scenes.stroke_scenes(30.0))  # the device stroker's kernels ran
    assert_same(g, o)
    # the persistent pen width is the same as after the host path (plotter.rs:151-153)
    monkeypatch.setenv("FTL_DEVICE_STROKE", "0")
    g2, _ = both(64, 64, Format.Matte8, join=join)
    for name, path in scenes.stroke_scenes(30.0).items():
        g2.stroke(path, (255,))
    assert g.pen_width() == g2.pen_width()


def test_device_stroker_hands_capped_strokes_to_the_host(monkeypatch):
    """A stroke at Stroke::add_point's 65 535-point cap (stroker.rs:206) is declined by the device stroker and drawn through
    the host stroker, in order, with the same pixels as the oracle."""
    monkeypatch.setenv("FTL_DEVICE_STROKE", "1")
    t = np.linspace(0.0, 40 * np.pi, 70000)
    r = 10 + 100 * t / t[-1]
    p = Path2D().absolute().pen_width(1.5).move_to(128 + r[0] * np.cos(t[0]), 128 + r[0] * np.sin(t[0]))
    for k in range(1, len(t)):
        p = p.line_to(128 + r[k] * np.cos(t[k]), 128 + r[k] * np.sin(t[k]))
    ops = p.finish()
    g, o = both(256, 256, Format.Matte8, join=JoinStyle.Bevel)
    assert g.debug_stroke_ops_device(ops) is None
    small = scenes.stroke_scenes(1.0)["round"]
    g.stroke(small, (255,))   # device
    g.stroke(ops, (255,))     # declined -> host
    g.stroke(small, (255,))   # device again, after the fallback
    for q in (small, ops, small):
        o.stroke(q, (255,))
    assert_same(g, o)


def test_batch_stroke_device_equals_host_path(monkeypatch):
    paths = list(scenes.stroke_scenes(4.0).values()) + [scenes.fishy_bench(), scenes.fishy_example()[0]]
    paths.insert(2, np.zeros(0, dtype=OP_DTYPE))                                 # a job without ops
    paths.insert(5, np.array([PathOp.PenWidth(4.0), PathOp.Close()], dtype=OP_DTYPE))  # a job without drawing ops
    ops, offs = Batch.pack(paths)
    n = len(paths)
    tr = np.tile(np.array([1, 0, 0, 0, 1, 0], dtype=np.float32), (n, 1))
    tr[1] = [0.8, 0.1, 6, -0.1, 1.1, 3]
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FTL_DEVICE_STROKE", mode)
        b = Batch(512, 512, Format.Matte8, n).set_join(JoinStyle.Round)
        b.stroke(ops, offs, transforms=tr)
        got[mode] = b.read()
    assert got["1"].any() and np.array_equal(got["1"], got["0"])


# ---- strict Vid(u16) mode: Fig::add_point's 65 535-point cap (fig.rs:428-442) -----------------------------------
@pytest.mark.parametrize("curved", [False, True])
def test_strict_vid_mode_reproduces_the_point_cap(curved):
    """ftl_set_strict_vid(1): points are ignored while 65 535 are stored, a Fig::close that pops a point makes room for
    one more (the oracle's default vid_cap is the reference's); without the flag every point is drawn (u32 ids)."""
    rng = np.random.default_rng(65535)
    p = Path2D().absolute()
    n_sub = 40 if curved else 90
    for s in range(n_sub):  # closed sub-figures of ~800 points each: the cap falls inside one of them
        cx, cy = rng.uniform(40, 216, 2)
        p = p.move_to(cx, cy)
        if curved:
            for k in range(60):
                p = p.quad_to(*(np.array([cx, cy, cx, cy]) + rng.uniform(-40, 40, 4)))
        else:
            for k in range(799):
                p = p.line_to(cx + rng.uniform(-30, 30), cy + rng.uniform(-30, 30))
        p = p.line_to(cx, cy).close()  # the closing point equals the first: popped by Fig::close
    ops = p.finish()
    tol = 0.01 if curved else None
    g, o = both(256, 256, Format.Matte8, tol=tol)
    g.set_strict_vid(True)
    g.fill(FillRule.EvenOdd, ops, (255,))
    o.fill(oracle.EVENODD, ops, (255,))
    assert 65500 <= o.last_info()["n_points"] <= 65535, o.last_info()  # the oracle did hit the cap (closing points popped since)
    assert_same(g, o, "strict")
    assert g.debug_last_fill() == o.last_info()
    strict = g.raster().pixels.copy()
    g2, o2 = both(256, 256, Format.Matte8, tol=tol, vid_cap=1 << 30)
    g2.fill(FillRule.EvenOdd, ops, (255,))
    o2.fill(oracle.EVENODD, ops, (255,))
    assert_same(g2, o2, "u32 ids")
    assert not np.array_equal(strict, g2.raster().pixels)
    # a fill below the cap is untouched by the flag
    small = scenes.fishy_example()[0]
    g3, o3 = both(128, 128, Format.Matte8)
    g3.set_strict_vid(True)
    g3.fill(FillRule.NonZero, small, (255,))
    o3.fill(oracle.NONZERO, small, (255,))
    assert_same(g3, o3, "below the cap")


def test_strict_vid_mode_a_popped_closing_point_makes_room():
    """65 533 distinct points and the closing point (= the first, popped by Fig::close, fig.rs:376-380) leave room for two
    more: of the triangle that follows only A and B are stored, the next triangle is ignored altogether."""
    n = 65533
    th = np.linspace(0.0, 2 * np.pi, n, endpoint=False)
    xs, ys = 128 + 100 * np.cos(th), 128 + 100 * np.sin(th)
    p = Path2D().absolute().move_to(xs[0], ys[0])
    for k in range(1, n):
        p = p.line_to(xs[k], ys[k])
    p = p.line_to(xs[0], ys[0]).close()
    p = p.move_to(100, 100).line_to(160, 110).line_to(120, 170).close()
    p = p.move_to(60, 60).line_to(90, 70).line_to(70, 95).close()
    ops = p.finish()
    g, o = both(256, 256, Format.Matte8)
    g.set_strict_vid(True)
    g.fill(FillRule.EvenOdd, ops, (255,))
    o.fill(oracle.EVENODD, ops, (255,))
    assert o.last_info()["n_points"] == 65535
    assert_same(g, o, "strict")
    assert g.debug_last_fill() == o.last_info()
    assert g.raster().pixels[120, 125] == 255  # inside the circle and inside the (undrawn) first triangle: EvenOdd would have cleared it
    g2, o2 = both(256, 256, Format.Matte8, vid_cap=1 << 30)
    g2.fill(FillRule.EvenOdd, ops, (255,))
    o2.fill(oracle.EVENODD, ops, (255,))
    assert_same(g2, o2, "u32 ids")
    assert g2.raster().pixels[120, 125] == 0


def test_profiling_state_is_per_handle():
    from footile_b200 import set_profiling, tile_kernel_time
    tile_kernel_time(reset=True)
    set_profiling(True)
    try:
        a = Plotter(Raster(256, 256, Format.Matte8))
        b = Plotter(Raster(256, 256, Format.Matte8))
        import os
        os.environ["FTL_NO_SMALL"] = "1"  # the one-launch small fill is not a tile-kernel launch
        try:
            for _ in range(3):
                a.fill(FillRule.NonZero, scenes.fishy_bench(), (255,))
            b.fill(FillRule.NonZero, scenes.fishy_bench(), (255,))
        finally:
            os.environ.pop("FTL_NO_SMALL", None)
        a.sync(); b.sync()
        ms_a, n_a = a.tile_kernel_time()
        ms_b, n_b = b.tile_kernel_time(reset=True)
        assert n_a == 3 * n_b and n_b >= 1 and ms_a > 0 and ms_b > 0
        ms_all, n_all = tile_kernel_time()
        assert n_all == n_a  # b was reset; the process-wide figure sums the live handles
        assert b.tile_kernel_time() == (0.0, 0)
    finally:
        set_profiling(False)
        tile_kernel_time(reset=True)


# ---- output conversion (examples/fishy.rs:33, examples/png/mod.rs:22-27) on the device ----------------------
@pytest.mark.parametrize("fmt", [Format.Rgba8p, Format.Graya8p, Format.Matte8])
def test_read_raster_srgb_matches_oracle_conversion(fmt):
    rng = np.random.default_rng(17)
    w, h = 256, 96
    bpp = {Format.Matte8: 1, Format.Graya8p: 2, Format.Rgba8p: 4}[fmt]
    px = rng.integers(0, 256, (h, w, bpp), dtype=np.uint8)
    if bpp > 1:  # premultiplied: colour <= alpha; include alpha 0 and 255 rows
        px[0, :, -1] = 0
        px[1, :, -1] = 255
        px[..., :-1] = np.minimum(px[..., :-1], px[..., -1:])
    init = px.reshape(h, w * bpp)
    g = Plotter(Raster(w, h, fmt, init))
    fish, eye = scenes.fishy_example()
    if fmt != Format.Matte8:
        g.fill(FillRule.NonZero, fish, (127, 96, 96, 255)[:bpp] if bpp == 4 else (127, 255))
    got = g.raster_srgb()
    exp = oracle.convert_srgb(FMT[fmt], g.raster().pixels)
    assert np.array_equal(got, exp)
    if fmt == Format.Matte8:
        assert np.array_equal(got, init)  # SGray8 view: the bytes themselves


def test_fishy_example_srgb_output_semantics():
    """examples/fishy.rs:29-33 end to end on the device; the conversion keeps alpha, leaves clear pixels clear and makes opaque
    mid-gray brighter (sRGB encode of linear 127/255 is 187)."""
    fish, eye = scenes.fishy_example()
    g = Plotter(Raster(128, 128, Format.Rgba8p))
    g.fill(FillRule.NonZero, fish, (127, 96, 96, 255))
    g.stroke(fish, (255, 208, 208, 255))
    g.stroke(eye, (0, 0, 0, 255))
    lin = g.raster().pixels.reshape(128, 128, 4)
    srgb = g.raster_srgb().reshape(128, 128, 4)
    assert np.array_equal(lin[..., 3], srgb[..., 3])
    assert not srgb[lin[..., 3] == 0].any()
    body = (lin[..., 3] == 255) & (lin[..., 0] == 127)
    assert body.any() and (srgb[body][:, 0] == 187).all()


# ---- stage (c) alone: the i16 area rows before the prefix sum -------------------------------------------------
def test_debug_area_rows_match_the_reference_area_buffer():
    """ftl_debug_area against the area buffer of the oracle's sequential scan (the reference's own loop, fig.rs:536-573),
    including a figure that starts above the raster (rows shift: SURVEY A.6-3) and spans that leave it on both sides."""
    rng = np.random.default_rng(31)
    for it in range(8):
        w, h = int(rng.integers(20, 200)), int(rng.integers(20, 120))
        n = int(rng.integers(3, 40))
        pts = [(float(rng.uniform(-30, w + 30)), float(rng.uniform(-25 if it & 1 else 2, h + 10))) for _ in range(n)]
        for general in (False, True):
            _general_path(general)
            try:
                g = Plotter(Raster(w, h, Format.Matte8))
                g.fill(FillRule.NonZero, poly(pts), (255,))
                _, info, area = oracle.fig_fill(w, h, oracle.MATTE8, 0, [pts], want_area=True)
                assert g.debug_last_fill()["top_row"] == info["top_row"]
                first = max(info["top_row"], 0)
                for row in sorted(set([first, min(first + 1, h - 1), h // 2, h - 1])):
                    if row < first:
                        continue
                    assert np.array_equal(g.debug_area(row), area[row]), (it, general, row)
            finally:
                _general_path(False)


# ---- several shards behind one C handle (ftl_ctx_*): a device index may repeat, so one GPU exercises the logic ------
def test_ctx_fill_batch_and_bands_equal_single_device_results():
    from footile_b200.sharding import Context
    ctx = Context([0, 0, 0])  # three shards on device 0
    assert ctx.size() == 3
    ops, offs, rules = scenes.random_curve_paths(500, 7, segments=12, size=160)
    got = ctx.fill_batch(160, 160, Format.Matte8, ops, offs, rules=rules)
    for j in range(7):
        o = oracle.Plotter(160, 160, oracle.MATTE8)
        o.fill(int(rules[j]), ops[int(offs[j]): int(offs[j + 1])], (255,))
        assert np.array_equal(got[j], o.raster()), j
    size = 700  # bands of 32-row multiples: 256 + 224 + 220 rows
    big = scenes.random_polygons(3, 90, vertices=24, size=size, extent=200)
    base = np.full((size, size * 4), 60, dtype=np.uint8)
    whole = ctx.fill_bands(size, size, Format.Rgba8p, FillRule.EvenOdd, big, (10, 200, 30, 180), init=base)
    o = oracle.Plotter(size, size, oracle.RGBA8P, init=base, vid_cap=1 << 30, orderfree=True)
    o.fill(oracle.EVENODD, big, (10, 200, 30, 180))
    assert np.array_equal(whole, o.raster())
