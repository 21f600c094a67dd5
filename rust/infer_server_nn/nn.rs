//! Drop-in replacement for `infer_server/src/nn.rs` of sgasse/infercam_onnx: same public items
//! (`Bbox`, `InferModel`, `UltrafaceVariant`, `UltrafaceModel::{new, run}`), tract and `image`
//! replaced by libultraface_b200.so. NOT COMPILED here (no Rust toolchain in the image).
use std::ffi::{CStr, CString};

use anyhow::{anyhow, Result};
use image::RgbImage;
use ultraface_sys as sys;

/// Bounding box defined as `[x_top_left, y_top_left, x_bottom_right, y_bottom_right]`.
pub type Bbox = [f32; 4];

pub trait InferModel {
    fn run(&self, input: &RgbImage) -> Result<Vec<(Bbox, f32)>>;
}

pub enum UltrafaceVariant {
    W640H480,
    W320H240,
}

impl UltrafaceVariant {
    pub fn width_height(&self) -> (u32, u32) {
        match self {
            UltrafaceVariant::W640H480 => (640, 480),
            UltrafaceVariant::W320H240 => (320, 240),
        }
    }
}

pub struct UltrafaceModel {
    handle: *mut sys::uf_model,
}

// The C library serialises calls on one handle internally (see include/ultraface_b200.h).
unsafe impl Send for UltrafaceModel {}
unsafe impl Sync for UltrafaceModel {}

fn last_error() -> anyhow::Error {
    let msg = unsafe { CStr::from_ptr(sys::uf_last_error()) }.to_string_lossy().into_owned();
    anyhow!(msg)
}

impl UltrafaceModel {
    pub async fn new(variant: UltrafaceVariant, max_iou: f32, min_confidence: f32) -> Result<Self> {
        let (width, height) = variant.width_height();
        // same cache location and file names as the original get_model(); the download step
        // (utils::download_file) stays in the caller's crate and is unchanged.
        let name = match variant {
            UltrafaceVariant::W640H480 => "ultraface-RFB-640.onnx",
            UltrafaceVariant::W320H240 => "ultraface-RFB-320.onnx",
        };
        let path = dirs::cache_dir().expect("cache dir").join("infercam_onnx").join(name);
        let c_path = CString::new(path.to_string_lossy().as_bytes())?;
        let mut handle = std::ptr::null_mut();
        let rc = unsafe {
            sys::uf_model_load(c_path.as_ptr(), width, height, max_iou, min_confidence, 0, 1, &mut handle)
        };
        if rc != sys::UF_OK {
            return Err(last_error());
        }
        Ok(Self { handle })
    }
}

impl InferModel for UltrafaceModel {
    fn run(&self, input: &RgbImage) -> Result<Vec<(Bbox, f32)>> {
        let mut dets = vec![sys::uf_det::default(); 256];
        let mut n: u32 = 0;
        loop {
            let rc = unsafe {
                sys::uf_infer(
                    self.handle,
                    input.as_raw().as_ptr(),
                    input.width(),
                    input.height(),
                    dets.as_mut_ptr(),
                    dets.len() as u32,
                    &mut n,
                )
            };
            if rc != sys::UF_OK {
                return Err(last_error());
            }
            if (n as usize) <= dets.len() {
                break;
            }
            dets.resize(n as usize, sys::uf_det::default());
        }
        Ok(dets[..n as usize].iter().map(|d| ([d.x0, d.y0, d.x1, d.y1], d.conf)).collect())
    }
}

impl Drop for UltrafaceModel {
    fn drop(&mut self) {
        unsafe { sys::uf_model_free(self.handle) }
    }
}
