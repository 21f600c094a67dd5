//! Drop-in replacement for `infer_server/src/nn.rs` of sgasse/infercam_onnx: same public items
//! (`Bbox`, `InferModel`, `UltrafaceVariant`, `UltrafaceModel::{new, run}`), tract and `image`
//! replaced by libultraface_b200.so. NOT COMPILED here (no Rust toolchain in the image).
//!
//! Knobs the reference hard-codes or does not have come from the environment so that the constructor signature
//! (nn.rs:55) is unchanged: `ULTRAFACE_DEVICE` (CUDA ordinal, default 0), `ULTRAFACE_MAX_BATCH` (default 64: what
//! `run_batch` accepts), `ULTRAFACE_LANES` (concurrent calls on the handle, default 2).
use std::ffi::{CStr, CString};

use anyhow::{anyhow, Result};
use image::RgbImage;
use ultraface_sys as sys;

use crate::utils::download_file;

const ULTRAFACE_LINK_640: &str = "https://github.com/onnx/models/raw/main/vision/body_analysis/ultraface/models/version-RFB-640.onnx";
const ULTRAFACE_LINK_320: &str = "https://github.com/onnx/models/raw/main/vision/body_analysis/ultraface/models/version-RFB-320.onnx";

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
    max_batch: u32,
}

// All entry points are thread-safe on one handle (include/ultraface_b200.h): `lanes` calls run concurrently,
// further callers wait their turn.
unsafe impl Send for UltrafaceModel {}
unsafe impl Sync for UltrafaceModel {}

pub(crate) fn last_error() -> anyhow::Error {
    let msg = unsafe { CStr::from_ptr(sys::uf_last_error()) }.to_string_lossy().into_owned();
    anyhow!(msg)
}

fn env_u32(name: &str, default: u32) -> u32 {
    std::env::var(name).ok().and_then(|v| v.parse().ok()).unwrap_or(default)
}

/// Where the weight file lives and where it comes from: the same cache directory, file names and URLs as the
/// original `get_model` (nn.rs:143-163), so an existing cache keeps working. Only the load that follows differs.
fn weight_file(variant: &UltrafaceVariant) -> (&'static str, &'static str) {
    match variant {
        UltrafaceVariant::W640H480 => ("ultraface-RFB-640.onnx", ULTRAFACE_LINK_640),
        UltrafaceVariant::W320H240 => ("ultraface-RFB-320.onnx", ULTRAFACE_LINK_320),
    }
}

pub(crate) async fn model_path(variant: &UltrafaceVariant) -> Result<std::path::PathBuf> {
    let (file_name, url) = weight_file(variant);
    let cache = dirs::cache_dir().ok_or_else(|| anyhow!("no cache directory on this platform"))?.join("infercam_onnx");
    std::fs::create_dir_all(&cache)?;
    let path = cache.join(file_name);
    if !path.is_file() {
        // first run: fetch the weights once, as the reference does
        download_file(&reqwest::Client::new(), url, &path).await?;
    }
    Ok(path)
}

pub(crate) fn model_config(path: &CString, variant: &UltrafaceVariant, max_iou: f32, min_confidence: f32) -> sys::uf_config {
    let (width, height) = variant.width_height();
    sys::uf_config {
        struct_size: std::mem::size_of::<sys::uf_config>() as u32,
        onnx_path: path.as_ptr(),
        net_w: width,
        net_h: height,
        max_iou,
        min_confidence,
        device: env_u32("ULTRAFACE_DEVICE", 0) as i32,
        max_batch: env_u32("ULTRAFACE_MAX_BATCH", 64),
        norm_preset: sys::UF_NORM_REFERENCE, // (x/255 - mean) / std, nn.rs:85-88
        chunk: 0,
        slots: 0,
        resize_round_intermediate: 0,
        flags: 0,
        lanes: env_u32("ULTRAFACE_LANES", 0),
        host_chunk: 0,
    }
}

impl UltrafaceModel {
    pub async fn new(variant: UltrafaceVariant, max_iou: f32, min_confidence: f32) -> Result<Self> {
        let path = model_path(&variant).await?;
        let c_path = CString::new(path.to_string_lossy().as_bytes())?;
        let cfg = model_config(&c_path, &variant, max_iou, min_confidence);
        let mut handle = std::ptr::null_mut();
        let rc = unsafe { sys::uf_model_load_ex(&cfg, &mut handle) };
        if rc != sys::UF_OK {
            return Err(last_error());
        }
        Ok(Self { handle, max_batch: cfg.max_batch })
    }

    /// Batched form of `run` (what a stream batcher calls): any frame sizes, results per frame.
    pub fn run_batch(&self, inputs: &[&RgbImage]) -> Result<Vec<Vec<(Bbox, f32)>>> {
        const CAP: usize = 256;
        let mut out = Vec::with_capacity(inputs.len());
        for group in inputs.chunks(self.max_batch as usize) {
            let ptrs: Vec<*const u8> = group.iter().map(|i| i.as_raw().as_ptr()).collect();
            let ws: Vec<u32> = group.iter().map(|i| i.width()).collect();
            let hs: Vec<u32> = group.iter().map(|i| i.height()).collect();
            let mut dets = vec![sys::uf_det::default(); group.len() * CAP];
            let mut counts = vec![0u32; group.len()];
            let rc = unsafe {
                sys::uf_infer_batch(self.handle, ptrs.as_ptr(), ws.as_ptr(), hs.as_ptr(), group.len() as u32,
                                    dets.as_mut_ptr(), CAP as u32, counts.as_mut_ptr())
            };
            if rc != sys::UF_OK {
                return Err(last_error());
            }
            for (i, &n) in counts.iter().enumerate() {
                let n = (n as usize).min(CAP);
                out.push(dets[i * CAP..i * CAP + n].iter().map(|d| ([d.x0, d.y0, d.x1, d.y1], d.conf)).collect());
            }
        }
        Ok(out)
    }
}

impl InferModel for UltrafaceModel {
    fn run(&self, input: &RgbImage) -> Result<Vec<(Bbox, f32)>> {
        let mut dets = vec![sys::uf_det::default(); 256];
        let mut n: u32 = 0;
        loop {
            let rc = unsafe {
                sys::uf_infer(self.handle, input.as_raw().as_ptr(), input.width(), input.height(), dets.as_mut_ptr(),
                              dets.len() as u32, &mut n)
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
