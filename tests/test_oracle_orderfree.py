"""The order-free per-(edge,row) closed form — the
formulation the CUDA kernels implement — must equal the sequential active-edge
scan of src/fig.rs:480-626 bit for bit.  Both live in the oracle; this test is
the evidence that replacing the active-edge list is exact."""
import numpy as np
import pytest

import oracle


def random_fig(rng, w, h):
    subs = []
    for _ in range(int(rng.integers(1, 4))):
        n = int(rng.integers(1, 10))
        snap = rng.random() < 0.4
        pts = []
        for _ in range(n):
            x = rng.uniform(-w, 2 * w)
            y = rng.uniform(-h / 2, 1.5 * h)
            if snap:
                x, y = round(x * 2) / 2, round(y * 2) / 2
            pts.append((x, y))
            if rng.random() < 0.15:
                pts.append((x, y))  # coincident point -> de-dup path
        if rng.random() < 0.3:
            pts.append(pts[0])  # explicit closing point -> pop path (fig.rs:376-380)
        subs.append(pts)
    return subs


@pytest.mark.parametrize("fmt", [oracle.MATTE8, oracle.RGBA8P, oracle.GRAYA8P])
def test_orderfree_equals_sequential(fmt):
    rng = np.random.default_rng(1234 + fmt)
    n_fail = 0
    for it in range(1500):
        w = int(rng.choice([1, 3, 8, 17, 40]))
        h = int(rng.choice([1, 3, 8, 17, 40]))
        subs = random_fig(rng, w, h)
        rule = int(rng.integers(0, 2))
        clr = rng.integers(0, 256, 4).astype(np.uint8)
        clr[:3] = np.minimum(clr[:3], clr[3])  # premultiplied
        if fmt == oracle.GRAYA8P:
            clr[0] = min(clr[0], clr[1])
        base = rng.integers(0, 256, (h, w * oracle.BPP[fmt])).astype(np.uint8)
        a, ia = oracle.fig_fill(w, h, fmt, rule, subs, clr=clr, raster=base, mode=0)
        b, ib = oracle.fig_fill(w, h, fmt, rule, subs, clr=clr, raster=base, mode=1)
        assert ia == ib
        if not np.array_equal(a, b):
            n_fail += 1
    assert n_fail == 0


def test_orderfree_negative_top_row_shift():
    # a figure whose top vertex is above the raster is drawn shifted down.
    subs = [[(1.0, -2.5), (6.0, 3.0), (0.5, 5.0)]]
    a, ia = oracle.fig_fill(8, 8, oracle.MATTE8, oracle.NONZERO, subs, mode=0)
    b, ib = oracle.fig_fill(8, 8, oracle.MATTE8, oracle.NONZERO, subs, mode=1)
    assert ia["top_row"] == -3 and ia == ib
    assert np.array_equal(a, b)
    assert a.any()
