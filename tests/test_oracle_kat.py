"""Pin the CPU oracle to the reference's own known-answer tests.

Every expected value below is the one asserted in the reference's in-tree unit
tests (cited per test); none is produced by this repo.
"""
import numpy as np
import pytest

import oracle
from oracle import Fixed as F

f, i, op = F.f, F.i, F.op


# ---- src/fixed.rs:166-313 -------------------------------------------------
def test_fixed_add():  # fixed.rs:166-174
    assert op("add", i(1), i(1)) == i(2)
    assert op("add", i(2), i(2)) == i(4)
    assert op("add", i(2), i(-2)) == i(0)
    assert op("add", i(2), i(-4)) == i(-2)
    assert op("add", f(1.5), f(1.5)) == i(3)
    assert op("add", f(3.5), f(-1.25)) == f(2.25)


def test_fixed_sub():  # fixed.rs:176-184
    assert op("sub", i(1), i(1)) == i(0)
    assert op("sub", i(3), i(2)) == i(1)
    assert op("sub", i(2), i(-2)) == i(4)
    assert op("sub", i(2), i(4)) == i(-2)
    assert op("sub", f(1.5), f(1.5)) == i(0)
    assert op("sub", f(3.5), f(1.25)) == f(2.25)


def test_fixed_mul():  # fixed.rs:186-194
    assert op("mul", i(2), i(2)) == i(4)
    assert op("mul", i(3), i(-2)) == i(-6)
    assert op("mul", i(4), f(0.5)) == i(2)
    assert op("mul", i(-16), i(-16)) == i(256)
    assert op("mul", i(37), i(3)) == i(111)
    assert op("mul", i(128), i(128)) == i(16384)


def test_fixed_div():  # fixed.rs:196-205
    assert op("div", i(4), i(2)) == i(2)
    assert op("div", i(-6), i(2)) == i(-3)
    assert op("div", i(2), f(0.5)) == i(4)
    assert op("div", i(256), i(-16)) == i(-16)
    assert op("div", i(111), i(3)) == i(37)
    assert op("div", i(37), i(3)) == f(12.33333)
    assert op("div", i(16384), i(128)) == i(128)


def test_fixed_shl():  # fixed.rs:207-214
    assert op("shl", i(0), 2) == i(0)
    assert op("shl", i(1), 1) == i(2)
    assert op("shl", f(0.5), 1) == i(1)
    assert op("shl", f(0.25), 2) == i(1)
    assert op("shl", f(0.125), 3) == i(1)


def test_fixed_shr():  # fixed.rs:216-223
    assert op("shr", i(0), 2) == i(0)
    assert op("shr", i(1), 1) == f(0.5)
    assert op("shr", i(2), 1) == i(1)
    assert op("shr", i(4), 2) == i(1)
    assert op("shr", i(8), 3) == i(1)


def test_fixed_abs():  # fixed.rs:225-232
    assert op("abs", i(1)) == i(1)
    assert op("abs", i(500)) == i(500)
    assert op("abs", i(-500)) == i(500)
    assert op("abs", f(-1.5)) == f(1.5)
    assert op("abs", f(-2.5)) == f(2.5)


def test_fixed_floor():  # fixed.rs:234-242
    assert op("floor", i(1)) == i(1)
    assert op("floor", i(500)) == i(500)
    assert op("floor", f(1.5)) == i(1)
    assert op("floor", f(1.99999)) == i(1)
    assert op("floor", f(-0.0001)) == i(-1)
    assert op("floor", f(-2.5)) == i(-3)


def test_fixed_ceil():  # fixed.rs:244-252
    assert op("ceil", i(1)) == i(1)
    assert op("ceil", i(500)) == i(500)
    assert op("ceil", f(1.5)) == i(2)
    assert op("ceil", f(1.99999)) == i(2)
    assert op("ceil", f(-0.0001)) == i(0)
    assert op("ceil", f(-2.5)) == i(-2)


def test_fixed_round():  # fixed.rs:254-264
    assert op("round", i(1)) == i(1)
    assert op("round", i(500)) == i(500)
    assert op("round", f(1.5)) == i(2)
    assert op("round", f(1.49999)) == i(1)
    assert op("round", f(1.99999)) == i(2)
    assert op("round", f(-0.0001)) == i(0)
    assert op("round", f(-2.5)) == i(-2)
    assert op("round", f(-2.9)) == i(-3)


def test_fixed_trunc():  # fixed.rs:266-276
    assert op("trunc", i(1)) == i(1)
    assert op("trunc", i(500)) == i(500)
    assert op("trunc", f(1.5)) == i(1)
    assert op("trunc", f(1.49999)) == i(1)
    assert op("trunc", f(1.99999)) == i(1)
    assert op("trunc", f(-0.0001)) == i(0)
    assert op("trunc", f(-2.5)) == i(-2)
    assert op("trunc", f(-2.9)) == i(-2)


def test_fixed_fract():  # fixed.rs:278-285
    assert op("fract", i(0)) == i(0)
    assert op("fract", f(0.1)) == f(0.1)
    assert op("fract", f(0.9)) == f(0.9)
    assert op("fract", f(1.5)) == f(0.5)
    assert op("fract", f(-2.5)) == f(0.5)


def test_fixed_avg():  # fixed.rs:287-294
    assert op("avg", i(1), i(2)) == f(1.5)
    assert op("avg", i(1), i(1)) == i(1)
    assert op("avg", i(5), i(-5)) == i(0)
    assert op("avg", i(3), i(37)) == i(20)
    assert op("avg", i(3), f(1.5)) == f(2.25)


def test_fixed_into():  # fixed.rs:296-304
    assert F.to_i32(i(37)) == 37
    assert F.to_f32(f(2.5)) == 2.5
    assert F.to_i32(f(2.5)) == 2


def test_fixed_cmp():  # fixed.rs:306-313
    assert i(37) > i(3)
    assert i(3) < i(37)
    assert i(-4) < i(4)
    assert min(i(37), i(3)) == i(3)
    assert max(i(37), i(3)) == i(37)


def test_fixed_from_f32_rust_cast_semantics():  # fixed.rs:88-93 with Rust `as i32` (saturating, NaN -> 0)
    assert f(float("nan")) == 0
    assert f(1e9) == 2**31 - 1
    assert f(-1e9) == -(2**31)
    assert f(-0.99999) == -65535  # truncation toward zero, not floor


# ---- src/fig.rs:691-794 ---------------------------------------------------
def test_fixed_pt():  # fig.rs:691-700
    a = (f(2.0), f(1.0))
    b = (f(3.0), f(4.0))
    c = (f(-1.0), f(1.0))
    assert (op("sub", b[0], a[0]), op("sub", b[1], a[1])) == (f(1.0), f(3.0))
    W = oracle.lib().orc_widdershins
    assert W(*a, *b) == 1
    assert W(*b, *a) == 0
    assert W(*b, *c) == 1


MODES = [pytest.param(0, True, id="seq-simd"), pytest.param(0, False, id="seq-scalar"), pytest.param(1, True, id="orderfree")]


@pytest.mark.parametrize("mode,simd", MODES)
def test_fig_3x3(mode, simd):  # fig.rs:702-721
    ras, _ = oracle.fig_fill(3, 3, oracle.RGBA8P, oracle.NONZERO, [[(1, 2), (1, 3), (2, 3), (2, 2)]],
                             clr=(99, 99, 99, 255), mode=mode, simd=simd)
    exp = np.zeros((3, 12), dtype=np.uint8)
    exp[2, 4:8] = (99, 99, 99, 255)
    assert np.array_equal(ras, exp)


@pytest.mark.parametrize("mode,simd", MODES)
def test_fig_9x1(mode, simd):  # fig.rs:723-735
    ras, _ = oracle.fig_fill(9, 1, oracle.MATTE8, oracle.NONZERO, [[(0, 0), (9, 1), (0, 1)]], mode=mode, simd=simd)
    assert ras.ravel().tolist() == [242, 213, 185, 156, 128, 100, 71, 43, 14]


@pytest.mark.parametrize("mode,simd", MODES)
def test_fig_x_bounds(mode, simd):  # fig.rs:737-749
    ras, _ = oracle.fig_fill(3, 3, oracle.MATTE8, oracle.NONZERO, [[(-1, 0), (-1, 3), (3, 1.5)]], mode=mode, simd=simd)
    assert ras.ravel().tolist() == [112, 16, 0, 255, 224, 32, 112, 16, 0]


@pytest.mark.parametrize("mode,simd", MODES)
def test_fig_partial(mode, simd):  # fig.rs:751-764
    ras, _ = oracle.fig_fill(1, 3, oracle.MATTE8, oracle.NONZERO, [[(0.5, 0), (0.5, 1.5), (1, 3), (1, 0)]], mode=mode, simd=simd)
    assert ras.ravel().tolist() == [128, 117, 43]


@pytest.mark.parametrize("mode,simd", MODES)
def test_fig_partial2(mode, simd):  # fig.rs:766-780
    ras, _ = oracle.fig_fill(3, 3, oracle.MATTE8, oracle.NONZERO, [[(1.5, 0), (1.5, 1.5), (2, 3), (3, 3), (3, 0)]],
                             mode=mode, simd=simd)
    assert ras.ravel().tolist() == [0, 128, 255, 0, 117, 255, 0, 43, 255]


@pytest.mark.parametrize("mode,simd", MODES)
def test_fig_partial3(mode, simd):  # fig.rs:782-794
    ras, _ = oracle.fig_fill(9, 1, oracle.MATTE8, oracle.NONZERO, [[(0, 0), (0, 0.3), (9, 0)]], mode=mode, simd=simd)
    assert ras.ravel().tolist() == [73, 64, 56, 47, 39, 30, 22, 13, 4]


# ---- src/imgbuf.rs:205-232 ------------------------------------------------
@pytest.mark.parametrize("simd", [True, False])
def test_accumulate_non_zero(simd):  # imgbuf.rs:205-221
    b = np.zeros(3000, dtype=np.int16)
    b[0] = 200
    a, z = oracle.accumulate(oracle.NONZERO, b, simd)
    assert (a == 200).all() and (z == 0).all()
    d = np.zeros(5000, dtype=np.int16)
    d[0] = 300
    c, z = oracle.accumulate(oracle.NONZERO, d, simd)
    assert (c == 255).all() and (z == 0).all()


@pytest.mark.parametrize("simd", [True, False])
def test_accumulate_even_odd(simd):  # imgbuf.rs:223-232
    b = np.zeros(3000, dtype=np.int16)
    b[0] = 300
    a, z = oracle.accumulate(oracle.EVENODD, b, simd)
    assert (a == 212).all() and (z == 0).all()


def test_accumulate_simd_equals_scalar_ragged():
    # widths not divisible by 8 (SURVEY §4 gap vii): SIMD body + scalar tail == scalar
    assert oracle.lib().orc_has_ssse3() == 1
    rng = np.random.default_rng(7)
    for n in (1, 7, 8, 9, 15, 63, 100, 1001):
        src = rng.integers(-700, 700, n).astype(np.int16)
        for rule in (oracle.NONZERO, oracle.EVENODD):
            a, _ = oracle.accumulate(rule, src, True)
            b, _ = oracle.accumulate(rule, src, False)
            assert np.array_equal(a, b)


# ---- src/plotter.rs:389-403 (smoke: must not crash) -----------------------
def test_plotter_overlapping_smoke():
    from footile_b200.path import Path2D
    path = (Path2D().absolute().move_to(8.0, 4.0).line_to(8.0, 3.0).cubic_to(8.0, 3.0, 8.0, 3.0, 9.0, 3.75)
            .line_to(8.0, 3.75).line_to(8.5, 3.75).line_to(8.5, 3.5).finish())
    for orderfree in (False, True):
        p = oracle.Plotter(16, 16, oracle.MATTE8, orderfree=orderfree)
        p.fill(oracle.NONZERO, path, (255,))
        assert p.raster().shape == (16, 16)


def test_batch_checksums_match_the_fold_of_ftl_batch_checksums():
    """orc_batch_fill_checksums (used by the full-size config 4 parity test) = 256 interleaved FNV-1a lanes folded in order."""
    from footile_b200 import scenes
    ops, offs, rules = scenes.random_curve_paths(7, 5, size=512)
    sums = oracle.batch_fill_checksums(512, 512, oracle.MATTE8, ops, offs, rules=rules, clr=(255,), threads=2)
    prime = np.uint64(0x100000001b3)
    for j in range(5):
        o = oracle.Plotter(512, 512, oracle.MATTE8)
        o.fill(int(rules[j]), ops[int(offs[j]): int(offs[j + 1])], (255,))
        lanes = np.ascontiguousarray(o.raster()).reshape(-1, 256).astype(np.uint64)
        h = np.full(256, 0xcbf29ce484222325, dtype=np.uint64)
        with np.errstate(over="ignore"):
            for row in lanes:
                h = (h ^ row) * prime
            g = np.uint64(0xcbf29ce484222325)
            for v in h:
                for k in range(8):
                    g = (g ^ ((v >> np.uint64(8 * k)) & np.uint64(0xFF))) * prime
        assert int(g) == int(sums[j]), j


def test_debug_edges_of_a_triangle():
    """Edge::new (fig.rs:179-210) on the fig_9x1 triangle (0,0)(9,1)(0,1): one shallow edge and one vertical one."""
    o = oracle.Plotter(9, 1, oracle.MATTE8)
    from footile_b200 import Path2D
    path = Path2D().absolute().move_to(0.0, 0.0).line_to(9.0, 1.0).line_to(0.0, 1.0).close().finish()
    e = o.debug_edges(path)
    assert e.shape == (2, 6)  # the horizontal side builds no edge (fig.rs:589: only edges going down)
    one = 1 << 16
    by_slope = {int(r[1]): r for r in e}
    shallow, vertical = by_slope[9 * one], by_slope[0]
    assert int(shallow[2]) == one // 9 and int(shallow[0]) == 9 * one and int(shallow[3]) == 0 and int(shallow[4]) == one
    assert int(vertical[2]) == 0 and int(vertical[0]) == 0 and int(vertical[3]) == 0 and int(vertical[4]) == one
    assert int(shallow[5]) == -int(vertical[5])  # the two sides wind opposite ways


def test_oracle_stroke_join_semantics_outer_corner():
    """The recalled pointy semantics (right(), angle_rel, Line::intersection: SURVEY App. B) cannot be checked
    against the crate, but they can be checked against GEOMETRY: examples/round.rs turns at (60,60) with the outer
    corner towards (80,80).  A flipped right() would put the join on the inner corner (stroker.rs:301-309,380-396)."""
    from footile_b200 import scenes
    path = scenes.stroke_scenes(1.0)["round"]
    imgs = {}
    for name, kind, limit in (("round", oracle.ROUND, 0.0), ("miter", oracle.MITER, 4.0), ("bevel", oracle.BEVEL, 0.0)):
        o = oracle.Plotter(100, 100, oracle.MATTE8)
        o.set_join(kind, limit)
        o.stroke(path, (255,))
        imgs[name] = o.raster()
    assert imgs["round"][73, 73] == 255 and imgs["round"][78, 78] == 0
    assert imgs["round"][66, 76] == 255 and imgs["round"][76, 66] == 255
    assert imgs["miter"][78, 78] == 255 and imgs["miter"][73, 73] == 255
    assert imgs["bevel"][73, 73] == 0 and imgs["bevel"][68, 68] == 255
    for img in imgs.values():
        assert img[45, 45] == 255 and img[35, 35] == 0
        assert img[70, 20] == 255 and img[20, 70] == 255 and img[5, 5] == 0
        assert img[60, 85] == 0 and img[85, 60] == 0


def test_oracle_stroke_width_and_coverage_against_geometry():
    """More guards on recalled semantics that geometry can check: the pen width is the FULL width of the
    stroke, split evenly to both sides (stroker.rs:301-309: offsets of w / 2); coverage integrates to the area; a closed
    path stroked with round joins covers its corners by discs of radius w / 2."""
    from footile_b200 import Path2D
    w, L = 10.0, 60.0
    line = Path2D().absolute().pen_width(w).move_to(20.25, 50.5).line_to(20.25 + L, 50.5).finish()
    o = oracle.Plotter(100, 100, oracle.MATTE8)
    o.stroke(line, (255,))
    img = o.raster().astype(np.float64) / 255.0
    assert abs(img.sum() - w * L) < 0.01 * w * L              # area of the w x L rectangle
    assert img[50, 50] == 1.0 and img[46, 50] == 1.0 and img[54, 50] == 1.0  # 5 px to either side of y = 50.5 ...
    assert img[44, 50] == 0.0 and img[56, 50] == 0.0                         # ... and nothing beyond
    assert abs(img[45, 50] - 0.5) < 0.01 and abs(img[55, 50] - 0.5) < 0.01   # the half-covered boundary rows (y = 45.5, 55.5)
    assert img[50, 19] == 0.0 and img[50, 81] == 0.0           # butt ends: nothing before x = 20.25 or after x = 80.25
    sq = Path2D().absolute().pen_width(8.0).move_to(30, 30).line_to(70, 30).line_to(70, 70).line_to(30, 70).close().finish()
    o = oracle.Plotter(100, 100, oracle.MATTE8)
    o.set_join(oracle.ROUND, 0.0)
    o.stroke(sq, (255,))
    img = o.raster()
    # NonZero fill of the outer and inner outlines: the ring between the 32 x 32 and 48 x 48 squares, corners rounded with radius 4
    assert img[50, 28] == 255 and img[50, 72] == 255 and img[28, 50] == 255 and img[50, 50] == 0
    assert img[27, 27] > 200 and img[26, 26] == 0 and img[26, 30] == 255  # corner disc about (30, 30), drawn as chords: (27.5, 27.5) is 3.5 away, (26.5, 26.5) is 4.9
    ring_area = 48 * 48 - 32 * 32 - (4 - np.pi) * 4.0 ** 2   # squares minus the four rounded-off corner pieces
    assert abs(img.astype(np.float64).sum() / 255.0 - ring_area) < 0.01 * ring_area


def test_oracle_flatten_stays_within_tolerance_of_the_curve():
    """plotter.rs:248-332 subdivides until the midpoint test passes: every flattened point lies ON the curve (midpoint
    subdivision evaluates the curve at dyadic parameters) and consecutive chords stay within ~tolerance of it."""
    from footile_b200 import Path2D
    a, b, c, d = (10.0, 200.0), (80.0, -150.0), (220.0, 420.0), (290.0, 60.0)
    ops = Path2D().absolute().move_to(*a).cubic_to(*b, *c, *d).finish()
    for tol in (0.3, 0.05):
        o = oracle.Plotter(300, 300, oracle.MATTE8)
        o.set_tolerance(tol)
        xy, subs = o.debug_flatten(ops)
        pts = xy.astype(np.float64) / 65536.0
        t = np.linspace(0.0, 1.0, 200001)[:, None]
        P = [np.array(p)[None, :] for p in (a, b, c, d)]
        curve = (1 - t) ** 3 * P[0] + 3 * (1 - t) ** 2 * t * P[1] + 3 * (1 - t) * t ** 2 * P[2] + t ** 3 * P[3]
        # every flattened point is (to Fixed / f32 rounding) a point of the curve
        for q in pts:
            assert np.min(np.hypot(curve[:, 0] - q[0], curve[:, 1] - q[1])) < 5e-3  # sample spacing of the dense curve is 3.5e-3
        # and the polyline deviates from the curve by no more than a small multiple of the tolerance
        seg_a, seg_b = pts[:-1], pts[1:]
        worst = 0.0
        for q in curve[::200]:
            ab = seg_b - seg_a
            u = np.clip(((q - seg_a) * ab).sum(1) / np.maximum((ab * ab).sum(1), 1e-12), 0.0, 1.0)
            worst = max(worst, np.min(np.hypot(*(seg_a + u[:, None] * ab - q).T)))
        assert worst < 2.0 * tol, (tol, worst)
        assert len(pts) > (20 if tol == 0.3 else 50)
