//! `footile_b200::Plotter` — footile's `Plotter` (src/plotter.rs:38-380) with the hot path
//! (flatten -> edge prep -> coverage -> accumulate / composite) on an NVIDIA B200 through the
//! C ABI of `libfootile_b200.so` (include/footile_b200.h).
//!
//! Switching is one line: `use footile_b200::Plotter;` instead of `use footile::Plotter;` —
//! `FillRule`, `PathOp`, `Path2D` and `JoinStyle` are re-exported from footile unchanged.
//!
//! ```ignore
//! use footile_b200::{FillRule, Path2D, Plotter};
//! use pix::{matte::Matte8, Raster};
//! let fish = Path2D::default().relative().pen_width(3.0).move_to(112.0, 24.0).line_to(-32.0, 24.0)
//!     .cubic_to(-96.0, -48.0, -96.0, 80.0, 0.0, 32.0).line_to(32.0, 24.0).line_to(-16.0, -40.0).close().finish();
//! let mut p = Plotter::new(Raster::with_clear(128, 128));
//! p.fill(FillRule::NonZero, &fish, Matte8::new(255));
//! ```
//!
//! STATUS: written against footile 0.7 / pix 0.14 / pointy 0.7 from their public API as used by the
//! reference's own sources and examples; NOT compiled in this repository's build image (no Rust
//! toolchain, crates not vendored).  The Python mirror `footile_b200/plotter.py` binds the same
//! symbols and is what the parity tests drive.
pub mod sys;

use std::borrow::Borrow;
use std::ffi::CStr;
use std::os::raw::c_void;

pub use footile::{FillRule, JoinStyle, Path2D, PathOp};
use pix::chan::{Ch8, Linear, Premultiplied};
use pix::el::Pixel;
use pix::Raster;
use pointy::{Pt, Transform};

use sys::*;

fn check(rc: i32) {
    // The reference has no Results: its failure modes are panics (fig.rs:489,492; plotter.rs:274).
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(ftl_last_error()) }.to_string_lossy().into_owned();
        panic!("footile_b200 (status {}): {}", rc, msg);
    }
}

/// PathOp (src/path.rs:18-31) -> `ftl_path_op`
fn lower(op: &PathOp) -> ftl_path_op {
    match *op {
        PathOp::Close() => ftl_path_op { tag: FTL_OP_CLOSE, v: [0.0; 6] },
        PathOp::Move(p) => ftl_path_op { tag: FTL_OP_MOVE, v: [p.x, p.y, 0.0, 0.0, 0.0, 0.0] },
        PathOp::Line(p) => ftl_path_op { tag: FTL_OP_LINE, v: [p.x, p.y, 0.0, 0.0, 0.0, 0.0] },
        PathOp::Quad(b, c) => ftl_path_op { tag: FTL_OP_QUAD, v: [b.x, b.y, c.x, c.y, 0.0, 0.0] },
        PathOp::Cubic(b, c, d) => ftl_path_op { tag: FTL_OP_CUBIC, v: [b.x, b.y, c.x, c.y, d.x, d.y] },
        PathOp::PenWidth(w) => ftl_path_op { tag: FTL_OP_PENWIDTH, v: [w, 0.0, 0.0, 0.0, 0.0, 0.0] },
    }
}

fn lower_all<T>(ops: T) -> Vec<ftl_path_op>
where
    T: IntoIterator,
    T::Item: Borrow<PathOp>,
{
    ops.into_iter().map(|o| lower(o.borrow())).collect()
}

/// The bytes of one pixel in raster order (what `ftl_fill` takes as `color`).
fn pixel_bytes<P: Pixel>(clr: P) -> Vec<u8> {
    Raster::<P>::with_color(1, 1, clr).as_u8_slice().to_vec()
}

/// Plotter for 2D vector paths (src/plotter.rs:22-56), rasterised on the GPU.
///
/// The raster lives in device memory while drawing; `raster()` / `raster_mut()` / `into_raster()`
/// bring it back (the only blocking points), so a sequence of `fill` / `stroke` calls runs
/// asynchronously on the handle's CUDA stream.
pub struct Plotter<P>
where
    P: Pixel<Chan = Ch8, Alpha = Premultiplied, Gamma = Linear>,
{
    handle: *mut ftl_plotter,
    raster: Raster<P>,
    /// the device copy is newer than `raster`
    device_newer: bool,
    /// `raster` was handed out mutably and may be newer than the device copy
    host_newer: bool,
}

impl<P> Plotter<P>
where
    P: Pixel<Chan = Ch8, Alpha = Premultiplied, Gamma = Linear>,
{
    /// Plotter::new (plotter.rs:96-115): takes ownership of the raster; its pixels are copied to HBM.
    pub fn new(raster: Raster<P>) -> Self {
        Self::new_on_device(raster, 0)
    }

    /// Extension: choose the CUDA device.
    pub fn new_on_device(raster: Raster<P>, device: i32) -> Self {
        let format = match std::mem::size_of::<P>() {
            1 => FTL_MATTE8,
            2 => FTL_GRAYA8P,
            4 => FTL_RGBA8P,
            n => panic!("footile_b200: unsupported pixel size {}", n),
        };
        let mut handle = std::ptr::null_mut();
        let px = raster.as_u8_slice();
        check(unsafe { ftl_plotter_new(raster.width(), raster.height(), format, px.as_ptr() as *const c_void, device, &mut handle) });
        Plotter { handle, raster, device_newer: false, host_newer: false }
    }

    /// plotter.rs:118-120
    pub fn width(&self) -> u32 {
        unsafe { ftl_width(self.handle) }
    }

    /// plotter.rs:123-125
    pub fn height(&self) -> u32 {
        unsafe { ftl_height(self.handle) }
    }

    /// plotter.rs:133-137 (clamped to >= 0.01 by the library, like the reference)
    pub fn set_tolerance(&mut self, t: f32) -> &mut Self {
        check(unsafe { ftl_set_tolerance(self.handle, t) });
        self
    }

    /// plotter.rs:140-143.  pointy's `Transform` does not expose its matrix, so it is recovered
    /// from three probe points; use `set_transform_matrix` to pass the six coefficients exactly.
    pub fn set_transform(&mut self, t: Transform<f32>) -> &mut Self {
        let o = t * Pt::new(0.0f32, 0.0);
        let x = t * Pt::new(1.0f32, 0.0);
        let y = t * Pt::new(0.0f32, 1.0);
        self.set_transform_matrix([x.x - o.x, y.x - o.x, o.x, x.y - o.y, y.y - o.y, o.y])
    }

    /// x' = e[0]*x + e[1]*y + e[2];  y' = e[3]*x + e[4]*y + e[5]
    pub fn set_transform_matrix(&mut self, e: [f32; 6]) -> &mut Self {
        check(unsafe { ftl_set_transform(self.handle, e.as_ptr()) });
        self
    }

    /// plotter.rs:158-161
    pub fn set_join(&mut self, js: JoinStyle) -> &mut Self {
        let (kind, limit) = match js {
            JoinStyle::Miter(ml) => (FTL_JOIN_MITER, ml),
            JoinStyle::Bevel => (FTL_JOIN_BEVEL, 0.0),
            JoinStyle::Round => (FTL_JOIN_ROUND, 0.0),
        };
        check(unsafe { ftl_set_join(self.handle, kind, limit) });
        self
    }

    /// Extension: reproduce the reference's 65 535-point cap per fill (fig.rs:430); off by default.
    pub fn set_strict_vid(&mut self, on: bool) -> &mut Self {
        check(unsafe { ftl_set_strict_vid(self.handle, on as i32) });
        self
    }

    fn push_host_pixels(&mut self) {
        if self.host_newer {
            let px = self.raster.as_u8_slice();
            check(unsafe { ftl_write_raster(self.handle, px.as_ptr() as *const c_void, px.len()) });
            self.host_newer = false;
        }
    }

    fn pull_device_pixels(&mut self) {
        if self.device_newer {
            let px = self.raster.as_u8_slice_mut();
            check(unsafe { ftl_read_raster(self.handle, px.as_mut_ptr() as *mut c_void, px.len()) });
            self.device_newer = false;
        }
    }

    /// Plotter::fill (plotter.rs:339-350).  Returns the plotter's raster like the reference; the
    /// pixels are fetched from the device here, so prefer `fill_async` when several calls follow
    /// each other and only the final raster matters.
    pub fn fill<T>(&mut self, rule: FillRule, ops: T, clr: P) -> &mut Raster<P>
    where
        T: IntoIterator,
        T::Item: Borrow<PathOp>,
    {
        self.fill_async(rule, ops, clr);
        self.raster_mut()
    }

    /// `fill` without the read-back: queued on the handle's stream.
    pub fn fill_async<T>(&mut self, rule: FillRule, ops: T, clr: P)
    where
        T: IntoIterator,
        T::Item: Borrow<PathOp>,
    {
        self.push_host_pixels();
        let v = lower_all(ops);
        let c = pixel_bytes(clr);
        let rule = match rule {
            FillRule::NonZero => 0,
            FillRule::EvenOdd => 1,
        };
        check(unsafe { ftl_fill(self.handle, rule, v.as_ptr(), v.len(), c.as_ptr()) });
        self.device_newer = true;
    }

    /// Plotter::stroke (plotter.rs:356-365)
    pub fn stroke<T>(&mut self, ops: T, clr: P) -> &mut Raster<P>
    where
        T: IntoIterator,
        T::Item: Borrow<PathOp>,
    {
        self.stroke_async(ops, clr);
        self.raster_mut()
    }

    /// `stroke` without the read-back.
    pub fn stroke_async<T>(&mut self, ops: T, clr: P)
    where
        T: IntoIterator,
        T::Item: Borrow<PathOp>,
    {
        self.push_host_pixels();
        let v = lower_all(ops);
        let c = pixel_bytes(clr);
        check(unsafe { ftl_stroke(self.handle, v.as_ptr(), v.len(), c.as_ptr()) });
        self.device_newer = true;
    }

    /// plotter.rs:368-370
    pub fn raster(&mut self) -> &Raster<P> {
        self.pull_device_pixels();
        &self.raster
    }

    /// plotter.rs:373-375: the caller may change pixels; they are sent back before the next drawing call.
    pub fn raster_mut(&mut self) -> &mut Raster<P> {
        self.pull_device_pixels();
        self.host_newer = true;
        &mut self.raster
    }

    /// plotter.rs:378-380
    pub fn into_raster(mut self) -> Raster<P> {
        self.pull_device_pixels();
        let raster = std::mem::replace(&mut self.raster, Raster::with_clear(0, 0));
        raster // Drop frees the device side
    }
}

impl<P> Drop for Plotter<P>
where
    P: Pixel<Chan = Ch8, Alpha = Premultiplied, Gamma = Linear>,
{
    fn drop(&mut self) {
        unsafe {
            ftl_plotter_free(self.handle);
        }
    }
}

// A handle is used from one thread at a time (`&mut self` on every drawing call, as in the reference);
// it may move between threads.
unsafe impl<P> Send for Plotter<P> where P: Pixel<Chan = Ch8, Alpha = Premultiplied, Gamma = Linear> {}
