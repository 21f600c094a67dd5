//! Raw bindings for `include/ultraface_b200.h` — every `UF_API` symbol, checked against the header by
//! `tests/test_host.py::test_rust_sys_crate_declares_every_header_symbol`. NOT COMPILED in this repository's
//! environment (no cargo/rustc in the image); kept mechanical so that review against the header suffices.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct uf_model {
    _private: [u8; 0],
}

#[repr(C)]
pub struct uf_batcher {
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

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct uf_glyph {
    pub x0: i32,
    pub y0: i32,
    pub w: u32,
    pub h: u32,
    pub offset: u32,
}

pub const UF_OK: c_int = 0;
pub const UF_ERR_INVALID_ARG: c_int = 1;
pub const UF_ERR_IO: c_int = 2;
pub const UF_ERR_ONNX: c_int = 3;
pub const UF_ERR_UNSUPPORTED: c_int = 4;
pub const UF_ERR_CUDA: c_int = 5;
pub const UF_ERR_NO_DEVICE: c_int = 6;
pub const UF_ERR_CAPACITY: c_int = 7;

pub const UF_NORM_REFERENCE: u32 = 0;
pub const UF_NORM_127_128: u32 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct uf_config {
    pub struct_size: u32,
    pub onnx_path: *const c_char,
    pub net_w: u32,
    pub net_h: u32,
    pub max_iou: f32,
    pub min_confidence: f32,
    pub device: i32,
    pub max_batch: u32,
    pub norm_preset: u32,
    pub chunk: u32,
    pub slots: u32,
    pub resize_round_intermediate: u32,
    pub flags: u32,
    pub lanes: u32,
    pub host_chunk: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct uf_info {
    pub net_w: u32,
    pub net_h: u32,
    pub num_priors: u32,
    pub num_layers: u32,
    pub num_tensors: u32,
    pub max_batch: u32,
    pub chunk: u32,
    pub slots: u32,
    pub weight_bytes: u64,
    pub workspace_bytes: u64,
    pub algorithmic_bytes_per_frame: u64,
    pub macs_per_frame: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct uf_kernel_stat {
    pub name: [c_char; 64],
    pub launches: u64,
    pub device_ms: f64,
    pub algorithmic_bytes: u64,
    pub compulsory_bytes: u64,
    pub flops: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct uf_jpeg_info {
    pub w: u32,
    pub h: u32,
    pub ncomp: u32,
    pub hs: [u32; 3],
    pub vs: [u32; 3],
    pub nblocks: u32,
    pub nonzero: u32,
    pub quant: [[u16; 64]; 3],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct uf_batcher_config {
    pub struct_size: u32,
    pub model: uf_config,
    pub devices: *const i32,
    pub n_devices: u32,
    pub max_batch: u32,
    pub max_delay_us: u32,
    pub capacity: u32,
    pub workers: u32,
    pub det_cap: u32,
    pub max_frame_bytes: u32,
    pub annotate_quality: u32,
    pub annotate_scale_w: f32,
    pub annotate_scale_h: f32,
    pub annotate_max_bytes: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct uf_result {
    pub stream: u64,
    pub user_tag: u64,
    pub device: i32,
    pub status: i32,
    pub n_dets: u32,
    pub batch_size: u32,
    pub latency_us: u64,
    pub file_bytes: u32,
    pub reserved_: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct uf_batcher_stats {
    pub submitted: u64,
    pub dropped: u64,
    pub completed: u64,
    pub failed: u64,
    pub batches: u64,
}

pub type uf_batch_fn = Option<
    unsafe extern "C" fn(
        user: *mut c_void,
        device: i32,
        rgb: *const *const u8,
        w: *const u32,
        h: *const u32,
        n: u32,
        out: *mut uf_det,
        cap: u32,
        n_out: *mut u32,
    ) -> c_int,
>;

extern "C" {
    // load / free
    pub fn uf_model_load(onnx_path: *const c_char, net_w: u32, net_h: u32, max_iou: f32, min_confidence: f32, device: i32,
                         max_batch: u32, out: *mut *mut uf_model) -> c_int;
    pub fn uf_model_load_ex(cfg: *const uf_config, out: *mut *mut uf_model) -> c_int;
    pub fn uf_model_free(m: *mut uf_model);
    pub fn uf_model_info(m: *const uf_model, out: *mut uf_info) -> c_int;
    // inference
    pub fn uf_infer(m: *mut uf_model, rgb: *const u8, w: u32, h: u32, out: *mut uf_det, cap: u32, n_out: *mut u32) -> c_int;
    pub fn uf_infer_batch(m: *mut uf_model, rgb: *const *const u8, w: *const u32, h: *const u32, n: u32, out: *mut uf_det,
                          cap: u32, n_out: *mut u32) -> c_int;
    pub fn uf_infer_batch_device(m: *mut uf_model, d_rgb: *const u8, w: u32, h: u32, n: u32, out: *mut uf_det, cap: u32,
                                 n_out: *mut u32) -> c_int;
    // parity hooks
    pub fn uf_raw_outputs(m: *mut uf_model, first: u32, n: u32, scores: *mut f32, boxes: *mut f32) -> c_int;
    pub fn uf_preproc_u8(m: *mut uf_model, rgb: *const u8, w: u32, h: u32, out_u8: *mut u8) -> c_int;
    pub fn uf_preproc_u8_batch(m: *mut uf_model, rgb: *const u8, w: u32, h: u32, n: u32, out_u8: *mut u8) -> c_int;
    pub fn uf_debug_prestem_u8(m: *mut uf_model, rgb: *const u8, w: u32, h: u32, n: u32, out_u8: *mut u8) -> c_int;
    pub fn uf_preproc_f32(m: *mut uf_model, rgb: *const u8, w: u32, h: u32, out: *mut f32) -> c_int;
    pub fn uf_postproc(m: *mut uf_model, scores: *const f32, boxes: *const f32, k: u32, out: *mut uf_det, cap: u32,
                       n_out: *mut u32, out_prior_idx: *mut i32) -> c_int;
    pub fn uf_tensor_count(m: *const uf_model, n: *mut u32) -> c_int;
    pub fn uf_tensor_info(m: *const uf_model, i: u32, onnx_name: *mut *const c_char, c: *mut u32, h: *mut u32, w: *mut u32) -> c_int;
    pub fn uf_tensor_read(m: *mut uf_model, i: u32, frame: u32, out_nchw: *mut f32) -> c_int;
    // pinned host memory
    pub fn uf_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn uf_host_free(p: *mut c_void);
    // measurement
    pub fn uf_profile_enable(m: *mut uf_model, on: c_int) -> c_int;
    pub fn uf_profile_reset(m: *mut uf_model) -> c_int;
    pub fn uf_profile_read(m: *mut uf_model, out: *mut uf_kernel_stat, cap: u32, n_out: *mut u32) -> c_int;
    pub fn uf_launch_count(m: *const uf_model, n: *mut u64) -> c_int;
    pub fn uf_debug_fail_after(m: *mut uf_model, stages: i32) -> c_int;
    // host-only
    pub fn uf_onnx_inspect(onnx_path: *const c_char, net_w: u32, net_h: u32, out: *mut c_char, cap: usize, needed: *mut usize) -> c_int;
    pub fn uf_resize_taps(src_len: u32, dst_len: u32, left: *mut i32, ntaps: *mut i32, w: *mut f32, w_pitch: u32,
                          max_taps: *mut u32) -> c_int;
    // JPEG in front of the path (N2)
    pub fn uf_infer_batch_jpeg(m: *mut uf_model, jpeg: *const *const u8, len: *const usize, n: u32, out: *mut uf_det, cap: u32,
                               n_out: *mut u32) -> c_int;
    pub fn uf_jpeg_coefficients_gpu(m: *mut uf_model, jpeg: *const u8, len: usize, coefs: *mut i16, cap_blocks: usize, on_device: *mut i32) -> c_int;
    pub fn uf_jpeg_decode_rgb(m: *mut uf_model, jpeg: *const u8, len: usize, out_rgb: *mut u8, cap_bytes: usize, w: *mut u32,
                              h: *mut u32) -> c_int;
    pub fn uf_jpeg_info_read(jpeg: *const u8, len: usize, out: *mut uf_jpeg_info) -> c_int;
    pub fn uf_jpeg_coefficients(jpeg: *const u8, len: usize, info: *mut uf_jpeg_info, coefs: *mut i16, cap_blocks: usize) -> c_int;
    // overlay + JPEG encode (N3, rectangles only)
    pub fn uf_annotate_encode_jpeg(m: *mut uf_model, rgb: *const u8, w: u32, h: u32, dets: *const uf_det, n_dets: u32, scale_w: f32,
                                   scale_h: f32, quality: u32, out: *mut u8, cap: usize, out_len: *mut usize) -> c_int;
    pub fn uf_annotate_reencode_jpeg(m: *mut uf_model, jpeg: *const u8, len: usize, dets: *const uf_det, n_dets: u32, scale_w: f32,
                                     scale_h: f32, quality: u32, out: *mut u8, cap: usize, out_len: *mut usize) -> c_int;
    pub fn uf_jpeg_write_coefficients(w: u32, h: u32, quality: u32, coefs: *const i16, n_blocks: usize, out: *mut u8, cap: usize,
                                      out_len: *mut usize) -> c_int;
    pub fn uf_jpeg_quality_tables(quality: u32, lum64: *mut u16, chr64: *mut u16) -> c_int;
    pub fn uf_annotate_reencode_batch_jpeg(m: *mut uf_model, jpeg: *const *const u8, len: *const usize, n: u32, dets: *const uf_det,
                                           det_counts: *const u32, scale_w: f32, scale_h: f32, quality: u32, out: *mut u8,
                                           out_stride: usize, out_len: *mut usize) -> c_int;
    pub fn uf_worker_batch_jpeg(m: *mut uf_model, jpeg: *const *const u8, len: *const usize, n: u32, scale_w: f32, scale_h: f32, quality: u32,
                                dets: *mut uf_det, cap: u32, n_dets: *mut u32, out: *mut u8, out_stride: usize, out_len: *mut usize) -> c_int;
    pub fn uf_text_atlas_set(m: *mut uf_model, charset: *const c_char, n_chars: u32, max_len: u32, glyphs: *const uf_glyph,
                             coverage: *const f32, n_coverage: usize) -> c_int;
    pub fn uf_confidence_text(confidence: f32, out: *mut c_char, cap: usize) -> c_int;
    pub fn uf_draw_boxes_rgb(m: *mut uf_model, rgb: *const u8, w: u32, h: u32, dets: *const uf_det, n_dets: u32, scale_w: f32,
                             scale_h: f32, out_rgb: *mut u8) -> c_int;
    // stream batcher + router
    pub fn uf_batcher_create(cfg: *const uf_batcher_config, out: *mut *mut uf_batcher) -> c_int;
    pub fn uf_batcher_create_ex(cfg: *const uf_batcher_config, f: uf_batch_fn, user: *mut c_void, out: *mut *mut uf_batcher) -> c_int;
    pub fn uf_batcher_destroy(b: *mut uf_batcher);
    pub fn uf_batcher_acquire(b: *mut uf_batcher, stream: u64, bytes: usize, buf: *mut *mut u8, ticket: *mut u64) -> c_int;
    pub fn uf_batcher_commit(b: *mut uf_batcher, ticket: u64, w: u32, h: u32, user_tag: u64) -> c_int;
    pub fn uf_batcher_abort(b: *mut uf_batcher, ticket: u64) -> c_int;
    pub fn uf_batcher_try_submit(b: *mut uf_batcher, stream: u64, rgb: *const u8, w: u32, h: u32, user_tag: u64,
                                 accepted: *mut i32) -> c_int;
    pub fn uf_batcher_commit_jpeg(b: *mut uf_batcher, ticket: u64, jpeg_len: usize, user_tag: u64) -> c_int;
    pub fn uf_batcher_try_submit_jpeg(b: *mut uf_batcher, stream: u64, jpeg: *const u8, len: usize, user_tag: u64,
                                      accepted: *mut i32) -> c_int;
    pub fn uf_batcher_ingest(b: *mut uf_batcher, msg: *const u8, len: usize, user_tag: u64, accepted: *mut i32, stream: *mut u64) -> c_int;
    pub fn uf_batcher_poll(b: *mut uf_batcher, res: *mut uf_result, dets: *mut uf_det, cap: u32, timeout_ms: u32,
                           n_out: *mut u32) -> c_int;
    pub fn uf_batcher_poll_frames(b: *mut uf_batcher, res: *mut uf_result, dets: *mut uf_det, files: *mut u8, file_stride: usize, cap: u32,
                                  timeout_ms: u32, n_out: *mut u32) -> c_int;
    pub fn uf_batcher_flush(b: *mut uf_batcher, timeout_ms: u32) -> c_int;
    pub fn uf_batcher_stats_read(b: *const uf_batcher, out: *mut uf_batcher_stats) -> c_int;
    pub fn uf_batcher_owner(b: *const uf_batcher, stream: u64, device: *mut i32) -> c_int;
    pub fn uf_batcher_model(b: *mut uf_batcher, device_slot: u32, out: *mut *mut uf_model) -> c_int;
    pub fn uf_debug_batcher_drive_msgs(b: *mut uf_batcher, msgs: *const *const u8, lens: *const usize, n_msgs: u32, total: u64,
                                       producers: u32, seconds: *mut f64, detections: *mut u64) -> c_int;
    pub fn uf_debug_batcher_drive(b: *mut uf_batcher, frames: *const u8, n_frames: u32, w: u32, h: u32, streams: *const u64,
                                  n_streams: u32, total: u64, producers: u32, seconds: *mut f64, detections: *mut u64) -> c_int;
    // ingest helpers
    pub fn uf_stream_hash(name: *const u8, len: usize, out: *mut u64) -> c_int;
    pub fn uf_protomsg_parse(msg: *const u8, len: usize, kind: *mut u32, id: *mut *const u8, id_len: *mut usize,
                             data: *mut *const u8, data_len: *mut usize) -> c_int;
    pub fn uf_debug_siphash(c: u32, d: u32, k0: u64, k1: u64, input: *const u8, len: usize, out: *mut u64) -> c_int;
    // misc
    pub fn uf_last_error() -> *const c_char;
    pub fn uf_version() -> *const c_char;
    pub fn uf_device_count(n: *mut i32) -> c_int;
}
