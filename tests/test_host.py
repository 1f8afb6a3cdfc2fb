"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the host-side path
vocabulary and stroker behave like the reference's, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import footile_b200 as fb
from footile_b200 import _lib, scenes
from footile_b200.path import OP_DTYPE, OpTag, Path2D, PathOp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "footile_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ftl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libfootile_b200.so does not export %s" % name
    assert declared == set(_lib.SYMBOLS)
    assert _lib.lib().ftl_abi_version() == 1


def test_no_cpu_fallback():
    if _has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(fb.FootileError) as e:
        fb.Plotter(fb.Raster(8, 8))
    assert e.value.status == 2  # FTL_ERR_NO_DEVICE
    with pytest.raises(fb.FootileError):
        fb.device_count()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "footile_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", "Makefile")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), "%s imports the oracle" % f
                assert not re.search(r"#include[^\n]*oracle|libfootile_oracle|orc_[a-z_]+\s*\(", src), "%s links the oracle" % f


def test_op_layout_matches_header():
    assert OP_DTYPE.itemsize == 28
    assert OP_DTYPE.fields["tag"][1] == 0 and OP_DTYPE.fields["v"][1] == 4
    assert [int(t) for t in OpTag] == [0, 1, 2, 3, 4, 5]


def test_path2d_relative_absolute_close():  # path.rs:64-171
    p = (Path2D().move_to(10, 10).line_to(5, 0).quad_to(1, 1, 2, 0).close().line_to(3, 4)
         .absolute().cubic_to(1, 2, 3, 4, 5, 6).pen_width(2.5).finish())
    assert [int(t) for t in p["tag"]] == [1, 2, 3, 0, 2, 4, 5]
    assert p["v"][1][:2].tolist() == [15, 10]          # relative to the pen
    assert p["v"][2][:4].tolist() == [16, 11, 17, 10]  # both points relative to the same pen
    assert p["v"][4][:2].tolist() == [3, 4]            # close() moved the builder pen to the origin
    assert p["v"][5][:6].tolist() == [1, 2, 3, 4, 5, 6]
    assert p["v"][6][0] == 2.5


def test_path2d_doc_example():  # lib.rs:10-27 / plotter.rs:23-37: builds without error
    fish, eye = scenes.fishy_example()
    assert len(fish) == 7 and len(eye) == 5
    assert fish["v"][3][:6].tolist() == [-16.0, 0.0, -16.0, 128.0, 80.0, 80.0]


def test_scene_generators_are_counter_based():
    a, offs, rules = scenes.random_curve_paths(10, 6)
    b, _, _ = scenes.random_curve_paths(12, 2)
    per = int(offs[1])
    assert a[2 * per: 4 * per].tobytes() == b.tobytes()      # sharding never changes a path
    assert rules.tolist() == [0, 1, 0, 1, 0, 1]
    v = a["v"][a["tag"] >= 1]
    assert v.min() >= 0 and v.max() < 512
    c = scenes.random_polygons(5, 3)
    d = scenes.random_polygons(6, 1)
    assert c[65:130].tobytes() == d.tobytes()
    assert c["v"][:, :2].min() >= 0 and c["v"][:, :2].max() < 32768


def _outline(join, limit, tol_sq, ops, counts, xyw):
    ops = fb.as_ops(ops)
    cap = 1 << 16
    out = np.zeros(cap, dtype=OP_DTYPE)
    n = C.c_size_t()
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    xyw = np.ascontiguousarray(xyw, dtype=np.float32)
    _lib.check(_lib.lib().ftl_debug_stroke_outline(join, limit, tol_sq, ops.ctypes.data, len(ops), counts.ctypes.data,
                                                   xyw.ctypes.data if xyw.size else None, out.ctypes.data, cap, C.byref(n)))
    assert n.value <= cap
    return out[: n.value]


@pytest.mark.parametrize("join,limit", [(0, 4.0), (0, 1.5), (0, 0.0), (1, 0.0), (2, 0.0)])
def test_host_stroker_matches_oracle(join, limit):
    """The product's host stroker (stroker.cpp) against the oracle's (stroker.rs restatement) on
    line-only paths, where flattening is the identity so no device is needed."""
    rng = np.random.default_rng(17 + join)
    paths = list(scenes.stroke_scenes(1.0).values())[:5]  # all but the curve scene
    for _ in range(40):
        p = Path2D().absolute().pen_width(float(rng.uniform(0.5, 12)))
        p = p.move_to(*rng.uniform(0, 100, 2))
        for _ in range(int(rng.integers(1, 9))):
            k = rng.random()
            if k < 0.7:
                p = p.line_to(*rng.uniform(0, 100, 2))
            elif k < 0.8:
                p = p.pen_width(float(rng.uniform(0.5, 12)))
            elif k < 0.9:
                p = p.close()
            else:
                p = p.move_to(*rng.uniform(0, 100, 2))
        paths.append(p.finish())
    for ops in paths:
        o = oracle.Plotter(8, 8, oracle.MATTE8)
        o.set_join(join, limit)
        exp = o.debug_stroke_ops(ops)
        wide = oracle.Plotter(8, 8, oracle.MATTE8).debug_flatten_wide(ops)
        counts = [1 if int(t) in (1, 2) else 0 for t in ops["tag"]]
        got = _outline(join, limit, np.float32(0.3) * np.float32(0.3), ops, counts, wide)
        assert len(got) == len(exp)
        assert got.tobytes() == exp.tobytes()


def test_ch8_mul_by_255_identity():
    """The device fast path for alpha = 0 / opaque pixels uses d*255 = d-1 for 1<=d<=15 else d
    (a consequence of pix's 12-bit Ch8 multiply); check it against the oracle's pix_compat exhaustively."""
    for d in range(256):
        got = int(oracle.src_over([d, 0], [0, 0], 0)[0])  # alpha 0: d' = 0 + d*(255-0)
        assert got == (d - 1 if 1 <= d <= 15 else d), d


def test_libm_restatement_matches_glibc():
    """csrc/libm_compat.cuh (what the device stroker evaluates for stroker.rs:305,354,389,407) returns the bits of this
    host's hypotf / atan2f: 4 x 2 M random inputs (pixel-scale, tie-heavy, mid-exponent and arbitrary bit patterns)."""
    from footile_b200.plotter import debug_libm_selftest
    h, a, s_wrong, s_undecided = debug_libm_selftest(2_000_000, 12345)
    assert (h, a, s_wrong) == (0, 0, 0)
    # the exhaustive sweep next to +-pi/2 (67 k floats) is decided everywhere; of the `>=` probes only those placed within
    # 1.2e-7 of their threshold may stay open (half of them are placed there on purpose)
    assert s_undecided <= 1_000_100


def test_rust_shim_binds_only_exported_symbols():
    """rust/footile-b200/src/sys.rs (the uncompiled Rust binding) declares nothing the library does not export,
    and its ftl_path_op mirrors the header's layout (tag + six floats)."""
    import ctypes
    import re
    from footile_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "rust", "footile-b200", "src", "sys.rs")).read()
    names = re.findall(r"pub fn (ftl_[a-z0-9_]+)\(", src)
    assert len(names) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert "pub tag: u32" in src and "pub v: [f32; 6]" in src


def test_strict_vid_intake_matches_the_oracle_point_list():
    """ftl_set_strict_vid's host-side replay of Fig::add_point / close with the 65 535-point cap (fig.rs:428-442,373-383): the
    points it keeps, once the closing points are popped, are the oracle's point list (its default vid_cap is the reference's)."""
    from footile_b200.plotter import debug_strict_intake
    rng = np.random.default_rng(4242)
    p = Path2D().absolute()
    for s in range(70):  # closed sub-figures of 1000 points + a closing point equal to the first: the cap falls inside one
        cx, cy = rng.uniform(40, 216, 2)
        p = p.move_to(cx, cy)
        for k in range(999):
            p = p.line_to(cx + rng.uniform(-30, 30), cy + rng.uniform(-30, 30))
        p = p.line_to(cx, cy).close()
    ops = p.finish()
    tr = [1.5, 0.25, 3.0, -0.5, 1.25, 7.0]
    kept = debug_strict_intake(ops, tr)
    assert kept is not None and set(kept["tag"]) == {int(OpTag.Move), int(OpTag.Line)}
    o = oracle.Plotter(256, 256, oracle.MATTE8)
    o.set_transform(tr)
    xy, subs = o.debug_flatten(ops)
    assert len(xy) <= 65535 and int(subs[:, 1].sum()) == len(xy)
    # our list: Fixed conversion (truncation towards zero, fixed.rs:88-93), one sub-figure per Move, closing points popped
    fx = np.trunc(kept["v"][:, :2].astype(np.float64) * 65536.0).astype(np.int64).astype(np.int32)
    starts = np.flatnonzero(kept["tag"] == int(OpTag.Move))
    ends = np.append(starts[1:], len(kept))
    got = []
    for a, b in zip(starts, ends):
        pts = fx[a:b]
        if len(pts) and (pts[-1] == pts[0]).all():
            pts = pts[:-1]
        if len(pts):
            got.append(pts)
    assert len(got) == len(subs)
    assert np.array_equal(np.concatenate(got), xy)
    assert [len(g) for g in got] == [int(n) for n in subs[:, 1]]
    # below the cap the intake stands aside
    assert debug_strict_intake(scenes.fishy_bench()) is None


def _stroke_subs_model(ops):
    """Stroke::add_point / close (stroker.rs:204-236) replayed op by op: every drawing op adds at least one point."""
    subs, have_points = [], False  # [first_op, end_op, joined, done]
    for i, op in enumerate(ops):
        tag = int(op["tag"])
        if tag == int(OpTag.PenWidth):
            continue
        if tag in (int(OpTag.Close), int(OpTag.Move)) and have_points:  # close(joined) acts on the CURRENT sub-stroke
            subs[-1][2] = 1 if tag == int(OpTag.Close) else 0
            subs[-1][3] = True
        if tag == int(OpTag.Close):
            continue
        if not subs or subs[-1][3]:  # add_point after done: a new sub-stroke
            subs.append([i, i + 1, 0, False])
        subs[-1][1] = i + 1
        have_points = True
    return [(a, b, j, 0) for a, b, j, _ in subs]


def test_stroke_sub_table_follows_stroke_add_point_and_close():
    """The op-level sub-stroke table of the device stroker against a direct replay of stroker.rs:204-236, including the
    reference's quirk that a Move after a Close un-joins the sub-stroke that Close had joined."""
    from footile_b200.plotter import debug_stroke_subs
    P = PathOp
    quirk = np.array([P.Move(1, 1), P.Line(5, 1), P.Line(5, 5), P.Close(), P.Move(9, 9), P.Line(12, 9), P.Close()], dtype=OP_DTYPE)
    assert [tuple(r) for r in debug_stroke_subs(quirk)] == [(0, 3, 0, 0), (4, 6, 1, 0)]
    rng = np.random.default_rng(31337)
    tags = [P.Close(), P.Move(1, 2), P.Line(3, 4), P.Quad(1, 2, 3, 4), P.Cubic(1, 2, 3, 4, 5, 6), P.PenWidth(2.0)]
    for it in range(300):
        n = int(rng.integers(0, 24))
        ops = np.array([tags[int(k)] for k in rng.choice(6, n, p=[0.2, 0.15, 0.3, 0.1, 0.1, 0.15])], dtype=OP_DTYPE) if n else np.zeros(0, dtype=OP_DTYPE)
        assert [tuple(int(v) for v in r) for r in debug_stroke_subs(ops)] == _stroke_subs_model(ops), it
