//! Batched replacement for `infer_server/src/inferer.rs` (the reference handles ONE frame per loop iteration in one task,
//! inferer.rs:29-50). NOT COMPILED here (no Rust toolchain in the image). The queue, the batching, the per-GPU routing and
//! the ordering live in libultraface_b200.so (`uf_batcher_*`); this file is the glue a maintainer adds:
//!
//! * `Inferer::submit` is what `FrameRouter::run` calls instead of `infer_tx.try_send_ref()` (router.rs:64-72): it decodes the
//!   JPEG straight INTO a pinned slot of the GPU that owns the stream (`uf_batcher_acquire`, N4) and commits it; like the
//!   reference's channel it never blocks and drops the frame when the queue is full.
//! * `Inferer::run` polls finished frames, draws, re-encodes and broadcasts exactly as before (inferer.rs:38-46).
use std::collections::HashMap;
use std::ffi::CString;
use std::sync::Mutex;

use image::RgbImage;
use ultraface_sys as sys;

use crate::nn::{last_error, model_config, model_path, Bbox, UltrafaceVariant};
use crate::{as_jpeg_stream_item, BroadcastSender};

const DET_CAP: usize = 64;

struct Pending {
    image: RgbImage, // kept for drawing; the pixels the GPU reads are the pinned copy
    sender: BroadcastSender,
    width: u32,
    height: u32,
}

pub struct Inferer {
    batcher: *mut sys::uf_batcher,
    pending: Mutex<HashMap<u64, Pending>>,
    next_tag: std::sync::atomic::AtomicU64,
}

unsafe impl Send for Inferer {}
unsafe impl Sync for Inferer {}

impl Inferer {
    pub async fn new() -> Self {
        let variant = UltrafaceVariant::W320H240; // inferer.rs:23
        let path = model_path(&variant).await.expect("failed to fetch model");
        let c_path = CString::new(path.to_string_lossy().as_bytes()).unwrap();
        let n_gpu = {
            let mut n = 0i32;
            unsafe { sys::uf_device_count(&mut n) };
            n.max(1)
        };
        let devices: Vec<i32> = (0..n_gpu).collect();
        let cfg = sys::uf_batcher_config {
            struct_size: std::mem::size_of::<sys::uf_batcher_config>() as u32,
            model: model_config(&c_path, &variant, 0.5, 0.5),
            devices: devices.as_ptr(),
            n_devices: devices.len() as u32,
            max_batch: 64,
            max_delay_us: 2000,
            capacity: 0,
            workers: 2,
            det_cap: DET_CAP as u32,
            max_frame_bytes: 1280 * 720 * 3,
            // 0: this glue draws and re-encodes on the CPU as the reference does. Set to 95 (and hand the frames over as JPEG with
            // `uf_batcher_ingest` / `uf_batcher_try_submit_jpeg`, polling with `uf_batcher_poll_frames`) and the library returns
            // the annotated, re-encoded frame itself: decode, draw and encode then run on the GPU too.
            annotate_quality: 0,
            annotate_scale_w: 0.0,
            annotate_scale_h: 0.0,
            annotate_max_bytes: 0,
        };
        let mut batcher = std::ptr::null_mut();
        let rc = unsafe { sys::uf_batcher_create(&cfg, &mut batcher) };
        assert!(rc == sys::UF_OK, "failed to initialize model: {}", last_error());
        Self { batcher, pending: Mutex::new(HashMap::new()), next_tag: 1.into() }
    }

    /// Called by the router with the stream key `hashed(&proto_msg.id)` (router.rs:58). Lossy.
    pub fn submit(&self, stream: u64, jpeg: &[u8], sender: BroadcastSender, width: u32, height: u32) -> bool {
        let header = match turbojpeg::read_header(jpeg) { Ok(h) => h, Err(_) => return false };
        let (w, h) = (header.width as u32, header.height as u32);
        let (mut buf, mut ticket) = (std::ptr::null_mut::<u8>(), 0u64);
        let rc = unsafe { sys::uf_batcher_acquire(self.batcher, stream, (w * h * 3) as usize, &mut buf, &mut ticket) };
        if rc != sys::UF_OK || buf.is_null() {
            return false; // queue full: dropped, like try_send_ref on a full channel
        }
        let pixels = unsafe { std::slice::from_raw_parts_mut(buf, (w * h * 3) as usize) };
        let out = turbojpeg::Image { pixels, width: w as usize, pitch: (w * 3) as usize, height: h as usize, format: turbojpeg::PixelFormat::RGB };
        if turbojpeg::Decompressor::new().and_then(|mut d| d.decompress(jpeg, out)).is_err() {
            unsafe { sys::uf_batcher_abort(self.batcher, ticket) };
            return false;
        }
        let image = RgbImage::from_raw(w, h, unsafe { std::slice::from_raw_parts(buf, (w * h * 3) as usize) }.to_vec()).unwrap();
        let tag = self.next_tag.fetch_add(1, std::sync::atomic::Ordering::Relaxed);
        self.pending.lock().unwrap().insert(tag, Pending { image, sender, width, height });
        unsafe { sys::uf_batcher_commit(self.batcher, ticket, w, h, tag) == sys::UF_OK }
    }

    /// Replaces the body of the original `run` loop after `infer_faces`: draw, encode, broadcast.
    pub fn run(&self) {
        let mut res = vec![sys::uf_result::default(); 256];
        let mut dets = vec![sys::uf_det::default(); 256 * DET_CAP];
        loop {
            let mut n = 0u32;
            let rc = unsafe { sys::uf_batcher_poll(self.batcher, res.as_mut_ptr(), dets.as_mut_ptr(), 256, 50, &mut n) };
            if rc != sys::UF_OK {
                continue;
            }
            for i in 0..n as usize {
                let r = res[i];
                let Some(p) = self.pending.lock().unwrap().remove(&r.user_tag) else { continue };
                if r.status != sys::UF_OK {
                    continue; // `if let Ok(..)`: a failed inference skips the frame (inferer.rs:37)
                }
                let k = (r.n_dets as usize).min(DET_CAP);
                let bboxes: Vec<(Bbox, f32)> =
                    dets[i * DET_CAP..i * DET_CAP + k].iter().map(|d| ([d.x0, d.y0, d.x1, d.y1], d.conf)).collect();
                let frame = crate::inferer_draw::draw_bboxes_on_image(p.image, bboxes, p.width, p.height);
                let buf = turbojpeg::compress_image(&frame, 95, turbojpeg::Subsamp::Sub2x2).expect("failed to compress");
                p.sender.send(as_jpeg_stream_item(&buf)).ok();
            }
        }
    }
}

impl Drop for Inferer {
    fn drop(&mut self) {
        unsafe { sys::uf_batcher_destroy(self.batcher) }
    }
}
