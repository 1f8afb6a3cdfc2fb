"""Path description: FillRule, PathOp, Path2D, JoinStyle.

Host-side mirror of footile's path vocabulary (reference: src/path.rs:9-171,
src/stroker.rs:13-20).  A path is a flat array of 28-byte ``ftl_path_op``
records (include/footile_b200.h) — the input format of the device pipeline.
"""
import enum

import numpy as np

# struct ftl_path_op { uint32_t tag; float v[6]; }
OP_DTYPE = np.dtype([("tag", "<u4"), ("v", "<f4", (6,))])
_F32 = np.float32


class FillRule(enum.IntEnum):
    """Fill-rule for filling paths (src/path.rs:9-14)."""
    NonZero = 0
    EvenOdd = 1


class OpTag(enum.IntEnum):
    """Discriminant of PathOp (src/path.rs:18-31)."""
    Close = 0
    Move = 1
    Line = 2
    Quad = 3
    Cubic = 4
    PenWidth = 5


class JoinStyle:
    """Style for stroke joins (src/stroker.rs:13-20)."""
    MITER, BEVEL, ROUND = 0, 1, 2

    def __init__(self, kind, limit=0.0):
        self.kind, self.limit = int(kind), float(limit)

    @classmethod
    def Miter(cls, limit):
        return cls(cls.MITER, limit)

    def __eq__(self, o):
        return isinstance(o, JoinStyle) and (self.kind, self.limit) == (o.kind, o.limit)

    def __repr__(self):
        return {0: "Miter(%g)" % self.limit, 1: "Bevel", 2: "Round"}[self.kind]


JoinStyle.Bevel = JoinStyle(JoinStyle.BEVEL)
JoinStyle.Round = JoinStyle(JoinStyle.ROUND)


class PathOp:
    """Constructors for single path operations (src/path.rs:18-31)."""

    @staticmethod
    def _mk(tag, *v):
        r = np.zeros((), dtype=OP_DTYPE)
        r["tag"] = int(tag)
        r["v"][: len(v)] = v
        return r

    @staticmethod
    def Close():
        return PathOp._mk(OpTag.Close)

    @staticmethod
    def Move(x, y):
        return PathOp._mk(OpTag.Move, x, y)

    @staticmethod
    def Line(x, y):
        return PathOp._mk(OpTag.Line, x, y)

    @staticmethod
    def Quad(bx, by, cx, cy):
        return PathOp._mk(OpTag.Quad, bx, by, cx, cy)

    @staticmethod
    def Cubic(bx, by, cx, cy, dx, dy):
        return PathOp._mk(OpTag.Cubic, bx, by, cx, cy, dx, dy)

    @staticmethod
    def PenWidth(w):
        return PathOp._mk(OpTag.PenWidth, w)


class Path2D:
    """Builder for an array of PathOp (src/path.rs:44-171).

    Relative coordinates (the default) are resolved at build time by adding
    the builder's pen in f32 (path.rs:80-86); ``close`` resets the builder pen
    to the origin (path.rs:89-93).
    """

    def __init__(self):
        self._ops = []
        self._absolute = False
        self._pen = (_F32(0), _F32(0))

    def absolute(self):
        self._absolute = True
        return self

    def relative(self):
        self._absolute = False
        return self

    def _pt(self, x, y):
        x, y = _F32(x), _F32(y)
        if self._absolute:
            return (x, y)
        return (_F32(self._pen[0] + x), _F32(self._pen[1] + y))

    def close(self):
        self._ops.append(PathOp.Close())
        self._pen = (_F32(0), _F32(0))
        return self

    def move_to(self, x, y):
        pb = self._pt(x, y)
        self._ops.append(PathOp.Move(*pb))
        self._pen = pb
        return self

    def line_to(self, x, y):
        pb = self._pt(x, y)
        self._ops.append(PathOp.Line(*pb))
        self._pen = pb
        return self

    def quad_to(self, bx, by, cx, cy):
        pb, pc = self._pt(bx, by), self._pt(cx, cy)
        self._ops.append(PathOp.Quad(*pb, *pc))
        self._pen = pc
        return self

    def cubic_to(self, bx, by, cx, cy, dx, dy):
        pb, pc, pd = self._pt(bx, by), self._pt(cx, cy), self._pt(dx, dy)
        self._ops.append(PathOp.Cubic(*pb, *pc, *pd))
        self._pen = pd
        return self

    def pen_width(self, width):
        self._ops.append(PathOp.PenWidth(width))
        return self

    def finish(self):
        """Finish path: returns the ops as a contiguous OP_DTYPE array."""
        if not self._ops:
            return np.zeros(0, dtype=OP_DTYPE)
        return np.array(self._ops, dtype=OP_DTYPE)


def as_ops(ops):
    """Coerce a Path2D result / list of PathOp / structured array to a contiguous OP_DTYPE array."""
    if isinstance(ops, Path2D):
        ops = ops.finish()
    if isinstance(ops, np.ndarray) and ops.dtype == OP_DTYPE:
        return np.ascontiguousarray(ops)
    ops = list(ops)
    if not ops:
        return np.zeros(0, dtype=OP_DTYPE)
    return np.ascontiguousarray(np.array(ops, dtype=OP_DTYPE))
