//! Raw declarations of the C ABI (include/footile_b200.h) that `Plotter` forwards to.
//! One line per entry point, in the header's order; the reference method each one replaces is
//! cited in the header.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

/// `ftl_path_op`: `{ uint32_t tag; float v[6]; }` (PathOp, src/path.rs:18-31)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ftl_path_op {
    pub tag: u32,
    pub v: [f32; 6],
}
pub const FTL_OP_CLOSE: u32 = 0;
pub const FTL_OP_MOVE: u32 = 1;
pub const FTL_OP_LINE: u32 = 2;
pub const FTL_OP_QUAD: u32 = 3;
pub const FTL_OP_CUBIC: u32 = 4;
pub const FTL_OP_PENWIDTH: u32 = 5;

pub const FTL_MATTE8: c_int = 0;
pub const FTL_GRAYA8P: c_int = 1;
pub const FTL_RGBA8P: c_int = 2;

pub const FTL_JOIN_MITER: c_int = 0;
pub const FTL_JOIN_BEVEL: c_int = 1;
pub const FTL_JOIN_ROUND: c_int = 2;

#[repr(C)]
pub struct ftl_plotter {
    _private: [u8; 0],
}
#[repr(C)]
pub struct ftl_batch {
    _private: [u8; 0],
}

extern "C" {
    pub fn ftl_abi_version() -> c_int;
    pub fn ftl_last_error() -> *const c_char;
    pub fn ftl_device_count(count: *mut c_int) -> c_int;

    pub fn ftl_plotter_new(w: u32, h: u32, format: c_int, init_pixels: *const c_void, device: c_int, out: *mut *mut ftl_plotter) -> c_int;
    pub fn ftl_plotter_new_band(w: u32, h: u32, row_begin: u32, row_end: u32, format: c_int, init_pixels: *const c_void, device: c_int,
                                out: *mut *mut ftl_plotter) -> c_int;
    pub fn ftl_plotter_free(p: *mut ftl_plotter) -> c_int;
    pub fn ftl_width(p: *const ftl_plotter) -> u32;
    pub fn ftl_height(p: *const ftl_plotter) -> u32;
    pub fn ftl_set_tolerance(p: *mut ftl_plotter, t: f32) -> c_int;
    pub fn ftl_set_transform(p: *mut ftl_plotter, e: *const f32) -> c_int;
    pub fn ftl_set_join(p: *mut ftl_plotter, join: c_int, miter_limit: f32) -> c_int;
    pub fn ftl_set_strict_vid(p: *mut ftl_plotter, enabled: c_int) -> c_int;
    pub fn ftl_pen_width(p: *const ftl_plotter) -> f32;
    pub fn ftl_fill(p: *mut ftl_plotter, rule: c_int, ops: *const ftl_path_op, n_ops: usize, color: *const u8) -> c_int;
    pub fn ftl_stroke(p: *mut ftl_plotter, ops: *const ftl_path_op, n_ops: usize, color: *const u8) -> c_int;
    pub fn ftl_fill_layers(p: *mut ftl_plotter, n_layers: u32, ops: *const ftl_path_op, op_offsets: *const u64, rules: *const u8,
                           colors: *const u8) -> c_int;
    pub fn ftl_stroke_outline(p: *mut ftl_plotter, ops: *const ftl_path_op, n_ops: usize, out: *mut ftl_path_op, cap: usize,
                              n_out: *mut usize) -> c_int;
    pub fn ftl_fill_upload(p: *mut ftl_plotter, rule: c_int, ops: *const ftl_path_op, n_ops: usize, color: *const u8) -> c_int;
    pub fn ftl_fill_replay(p: *mut ftl_plotter) -> c_int;
    pub fn ftl_read_raster(p: *mut ftl_plotter, dst: *mut c_void, nbytes: usize) -> c_int;
    pub fn ftl_read_raster_srgb(p: *mut ftl_plotter, dst: *mut c_void, nbytes: usize) -> c_int;
    pub fn ftl_write_raster(p: *mut ftl_plotter, src: *const c_void, nbytes: usize) -> c_int;
    pub fn ftl_sync(p: *mut ftl_plotter) -> c_int;

    pub fn ftl_batch_new(w: u32, h: u32, format: c_int, capacity: u32, device: c_int, out: *mut *mut ftl_batch) -> c_int;
    pub fn ftl_batch_free(b: *mut ftl_batch) -> c_int;
    pub fn ftl_batch_set_tolerance(b: *mut ftl_batch, t: f32) -> c_int;
    pub fn ftl_batch_set_join(b: *mut ftl_batch, join: c_int, miter_limit: f32) -> c_int;
    pub fn ftl_batch_clear(b: *mut ftl_batch, first: u32, count: u32) -> c_int;
    pub fn ftl_batch_fill(b: *mut ftl_batch, n_jobs: u32, ops: *const ftl_path_op, op_offsets: *const u64, rules: *const u8,
                          transforms: *const f32, colors: *const u8) -> c_int;
    pub fn ftl_batch_stroke(b: *mut ftl_batch, n_jobs: u32, ops: *const ftl_path_op, op_offsets: *const u64, transforms: *const f32,
                            colors: *const u8) -> c_int;
    pub fn ftl_batch_read(b: *mut ftl_batch, first: u32, count: u32, dst: *mut c_void, nbytes: usize) -> c_int;
    pub fn ftl_batch_sync(b: *mut ftl_batch) -> c_int;
}
