//! Raw bindings for `include/ultraface_b200.h`. NOT COMPILED in this repository's environment
//! (no cargo/rustc in the image); kept thin so that review against the header suffices.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct uf_model {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct uf_det {
    pub x0: f32,
    pub y0: f32,
    pub x1: f32,
    pub y1: f32,
    pub conf: f32,
}

pub const UF_OK: c_int = 0;

extern "C" {
    pub fn uf_model_load(
        onnx_path: *const c_char,
        net_w: u32,
        net_h: u32,
        max_iou: f32,
        min_confidence: f32,
        device: i32,
        max_batch: u32,
        out: *mut *mut uf_model,
    ) -> c_int;
    pub fn uf_model_free(m: *mut uf_model);
    pub fn uf_infer(
        m: *mut uf_model,
        rgb: *const u8,
        w: u32,
        h: u32,
        out: *mut uf_det,
        cap: u32,
        n_out: *mut u32,
    ) -> c_int;
    pub fn uf_infer_batch(
        m: *mut uf_model,
        rgb: *const *const u8,
        w: *const u32,
        h: *const u32,
        n: u32,
        out: *mut uf_det,
        cap: u32,
        n_out: *mut u32,
    ) -> c_int;
    pub fn uf_last_error() -> *const c_char;
}
