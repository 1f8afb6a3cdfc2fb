// pointy_compat.cuh — the f32 point/transform arithmetic the hot path takes
// from the `pointy 0.7` crate (Cargo.toml:16).  Its source is not in the
// reference tree, so the op order below is RECALLED and
// kept in this one header so a later check against the real crate changes one
// place.  Compile with -fmad=false: Rust never contracts a*b+c into an FMA.
#pragma once
#include "fixed.cuh"

namespace ftl {
namespace pointy {

struct Pt { float x, y; };

FTL_HD Pt add(Pt a, Pt b) { return {a.x + b.x, a.y + b.y}; }
FTL_HD Pt sub(Pt a, Pt b) { return {a.x - b.x, a.y - b.y}; }
FTL_HD Pt scale(Pt a, float s) { return {a.x * s, a.y * s}; }
// Pt::midpoint (call sites: geom.rs:32, stroker.rs:409)
FTL_HD Pt midpoint(Pt a, Pt b) { return {(a.x + b.x) / 2.0f, (a.y + b.y) / 2.0f}; }
// Pt::distance_sq (call sites: plotter.rs:275, stroker.rs:129)
FTL_HD float distance_sq(Pt a, Pt b) {
    float dx = a.x - b.x, dy = a.y - b.y;
    return dx * dx + dy * dy;
}
// Transform * Pt (call site: plotter.rs:170); e = [a b tx; c d ty]
FTL_HD Pt transform(const float *e, Pt p) {
    return {e[0] * p.x + e[1] * p.y + e[2], e[3] * p.x + e[4] * p.y + e[5]};
}
FTL_HD Pt right(Pt v) { return {v.y, -v.x}; }
FTL_HD float cross(Pt a, Pt b) { return a.x * b.y - a.y * b.x; }

}  // namespace pointy
}  // namespace ftl
