/*
 * ultraface_b200.h — C ABI of the B200-native face-detection hot path.
 *
 * Drop-in boundary for sgasse/infercam_onnx's `infer_server::nn` module
 * (citations: /root/reference/infer_server/src/nn.rs). A Rust `ultraface-sys`
 * crate binds exactly these symbols (see INTEGRATION.md and rust/); the Python
 * mirror used by the tests is infercam_onnx_b200/nn.py.
 *
 * Conventions: every function returns UF_OK (0) or a uf_status error code and
 * never aborts the process; uf_last_error() returns a thread-local message for
 * the last failing call on this thread. All entry points are thread-safe on one
 * handle (`UltrafaceModel` must be Send + Sync: it is borrowed across .await in
 * inferer.rs:29-50). A handle owns `lanes` independent pipelines: up to that many
 * calls from different host threads run concurrently (the kernels of one batch
 * overlap the host-to-device copies of the next), further callers wait their turn.
 * The parity hooks (uf_raw_outputs, uf_tensor_read) refer to the last batch of
 * lane 0, which is the lane a single-threaded caller always gets.
 * No torch types, plain pointers and sizes only.
 */
#ifndef ULTRAFACE_B200_H
#define ULTRAFACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UF_ABI_VERSION 1

#if defined(__GNUC__)
#define UF_API __attribute__((visibility("default")))
#else
#define UF_API
#endif

typedef enum uf_status {
    UF_OK = 0,
    UF_ERR_INVALID_ARG = 1,
    UF_ERR_IO = 2,          /* model file missing / unreadable (nn.rs:156-162 would download) */
    UF_ERR_ONNX = 3,        /* malformed protobuf / missing initialiser */
    UF_ERR_UNSUPPORTED = 4, /* graph uses an operator/shape outside the UltraFace family */
    UF_ERR_CUDA = 5,        /* CUDA runtime error (message carries cudaGetErrorString) */
    UF_ERR_NO_DEVICE = 6,   /* no CUDA device: there is NO CPU fallback */
    UF_ERR_CAPACITY = 7     /* batch larger than max_batch, or K larger than the model's */
} uf_status;

/* Opaque model handle. Replaces `struct UltrafaceModel` (nn.rs:45-51). */
typedef struct uf_model uf_model;

/*
 * One detection. Replaces the element type of `Vec<(Bbox, f32)>` (nn.rs:12,25):
 * Bbox = [x_top_left, y_top_left, x_bottom_right, y_bottom_right], relative
 * coordinates (unclamped), followed by the confidence. Rust tuples have no
 * stable layout, so the wrapper converts uf_det -> (Bbox, f32).
 */
typedef struct uf_det {
    float x0, y0, x1, y1;
    float conf;
} uf_det;

/* Normalisation presets for nn.rs:82-91. */
#define UF_NORM_REFERENCE 0u /* (x/255 - mean[c]) / std[c], ImageNet constants (nn.rs:86-88) */
#define UF_NORM_127_128 1u   /* (x - 127) / 128, the upstream UltraFace convention */

typedef struct uf_config {
    uint32_t struct_size;    /* = sizeof(uf_config) */
    const char* onnx_path;   /* UltraFace ONNX file (nn.rs:145-157 cache path); required */
    uint32_t net_w, net_h;   /* UltrafaceVariant::width_height (nn.rs:36-41): 640x480 / 320x240 */
    float max_iou;           /* nn.rs:55 */
    float min_confidence;    /* nn.rs:55 */
    int32_t device;          /* CUDA device ordinal */
    uint32_t max_batch;      /* largest n accepted by uf_infer_batch* (>=1) */
    uint32_t norm_preset;    /* UF_NORM_* */
    uint32_t chunk;          /* frames per launch for device-resident input; 0 = auto */
    uint32_t slots;          /* pipeline depth (streams); 0 = auto */
    uint32_t resize_round_intermediate; /* 0 = image 0.24.x (f32 between passes); 1 = pre-0.24 */
    uint32_t flags;          /* UF_FLAG_* */
    uint32_t lanes;          /* concurrent calls served in parallel on one handle; 0 = auto (2) */
    uint32_t host_chunk;     /* frames per pipeline stage for HOST input (H2D overlap); 0 = auto */
} uf_config;

#define UF_FLAG_FORCE_GENERIC 1u /* debug: run every conv through the generic direct kernel */
#define UF_FLAG_NO_GRAPH 2u      /* do not capture CUDA graphs for the batch-1 path */
#define UF_FLAG_NO_FUSION 4u     /* debug: keep depthwise and pointwise convs as separate kernels */
#define UF_FLAG_NO_TC 8u         /* keep the 1x1 convs on the fp32 SIMT kernels (no tcgen05 3xTF32 path) */
#define UF_FLAG_PDL 32u          /* launch the kernel chain with programmatic dependent launch (process-wide; measured neutral) */
#define UF_FLAG_TMA_SIMT_PW 64u  /* fused dw+1x1 layers of the big maps: keep the 1x1 on the SIMT pipes (the pre-tcgen05 kernel) */
#define UF_FLAG_DENSE3_TC 128u   /* dense 3x3 convs of the RFB branches as a tcgen05 implicit GEMM (zero-copy im2col; measured slower) */
#define UF_FLAG_JPEG_HOST_HUFFMAN 512u /* uf_infer_batch_jpeg: Huffman-decode on host threads (default: on the GPU, host only for restart intervals / damaged streams) */
#define UF_FLAG_NO_PRESTEM 256u  /* frames at exactly 2x the network size: keep resize and stem as two kernels (default: one fused kernel) */
#define UF_FLAG_FUSE_DW_TC 16u   /* compute depthwise 3x3 inside the tensor-core GEMM's converter warps (C >= 64) */

typedef struct uf_info {
    uint32_t net_w, net_h;
    uint32_t num_priors;     /* K: 4420 @320x240, 17640 @640x480 */
    uint32_t num_layers;     /* device conv launches per chunk (after fusion) */
    uint32_t num_tensors;    /* materialised activation tensors (uf_tensor_*) */
    uint32_t max_batch, chunk, slots;
    uint64_t weight_bytes;
    uint64_t workspace_bytes;
    uint64_t algorithmic_bytes_per_frame; /* SURVEY.md §8(d) figure for this graph, from 640x480 input */
    uint64_t macs_per_frame;
} uf_info;

/* ---- load / free: replaces UltrafaceModel::new + get_model (nn.rs:55-67,143-175) ---- */
UF_API int uf_model_load(const char* onnx_path, uint32_t net_w, uint32_t net_h, float max_iou,
                  float min_confidence, int32_t device, uint32_t max_batch, uf_model** out);
UF_API int uf_model_load_ex(const uf_config* cfg, uf_model** out);
UF_API void uf_model_free(uf_model* m);
UF_API int uf_model_info(const uf_model* m, uf_info* out);

/* ---- per-frame call: replaces `InferModel::run(&self, &RgbImage)` (nn.rs:24-26,178-186) ----
 * rgb: contiguous HWC u8, w*h*3 bytes, any size (RgbImage::as_raw()).
 * out: up to cap detections in descending confidence (selection order, nn.rs:134-137);
 * *n_out = number selected (if > cap only the first cap were written). */
UF_API int uf_infer(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uf_det* out, uint32_t cap,
             uint32_t* n_out);

/* ---- batched call (what a stream batcher in inferer.rs:29-50 would call) ----
 * rgb[i]: host pointer to frame i (pinned or pageable), w[i] x h[i]; out: n x cap; n_out: n. */
UF_API int uf_infer_batch(uf_model* m, const uint8_t* const* rgb, const uint32_t* w, const uint32_t* h,
                   uint32_t n, uf_det* out, uint32_t cap, uint32_t* n_out);

/* Same, n frames of identical size contiguous in DEVICE memory (HBM-resident input). */
UF_API int uf_infer_batch_device(uf_model* m, const uint8_t* d_rgb, uint32_t w, uint32_t h, uint32_t n,
                          uf_det* out, uint32_t cap, uint32_t* n_out);

/* ---- parity hooks (host buffers) ---- */
/* Raw network outputs of frames [first, first+n) of the last batch: tract's outputs[0]
 * `scores` n x K x 2 and outputs[1] `boxes` n x K x 4 (nn.rs:111-120). */
UF_API int uf_raw_outputs(uf_model* m, uint32_t first, uint32_t n, float* scores, float* boxes);
/* nn.rs:74-80 alone: out_u8 = net_h x net_w x 3. */
UF_API int uf_preproc_u8(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t* out_u8);
/* Same for n <= chunk frames of identical size, resized by ONE launch (the batch path's kernel selection, e.g. the
 * exact-integer 2:1 kernel, differs from the single-frame one): out_u8 = n x net_h x net_w x 3. */
UF_API int uf_preproc_u8_batch(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t n, uint8_t* out_u8);
/* The resized u8 pixels as the FUSED resize + normalise + stem kernel (frames at exactly 2x the network size) computed and
 * convolved them: n <= chunk frames of (2 net_w) x (2 net_h) -> out_u8 = n x net_h x net_w x 3. UF_ERR_UNSUPPORTED otherwise. */
UF_API int uf_debug_prestem_u8(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t n, uint8_t* out_u8);
/* nn.rs:70-94: out = 1 x 3 x net_h x net_w f32 (NCHW). */
UF_API int uf_preproc_f32(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, float* out);
/* nn.rs:109-140 + 198-260 alone on caller-supplied raw tensors (scores K x 2, boxes K x 4). */
UF_API int uf_postproc(uf_model* m, const float* scores, const float* boxes, uint32_t K, uf_det* out,
                uint32_t cap, uint32_t* n_out, int32_t* out_prior_idx /* nullable, cap */);
/* Materialised activation tensors of the last batch, addressed by ONNX value name. */
UF_API int uf_tensor_count(const uf_model* m, uint32_t* n);
UF_API int uf_tensor_info(const uf_model* m, uint32_t i, const char** onnx_name, uint32_t* c, uint32_t* h,
                   uint32_t* w);
/* Copies tensor i of frame `frame` of the last batch's FIRST chunk as NCHW f32 (c*h*w floats). */
UF_API int uf_tensor_read(uf_model* m, uint32_t i, uint32_t frame, float* out_nchw);

/* ---- pinned host memory for the frames fed to uf_infer_batch ---- */
UF_API int uf_host_alloc(size_t bytes, void** out);
UF_API void uf_host_free(void* p);

/* ---- measurement: per-kernel-family CUDA-event timing on the launching streams ---- */
typedef struct uf_kernel_stat {
    char name[64];              /* "<kernel family>[<cin>><cout> k s d WxH]" */
    uint64_t launches;
    double device_ms;          /* sum of CUDA-event durations */
    uint64_t algorithmic_bytes; /* SURVEY.md 8(d) convention: sum over Conv nodes of (in+out)*4 */
    uint64_t compulsory_bytes;  /* what the launch must move (fusion removes the intermediate) */
    uint64_t flops;
} uf_kernel_stat;
UF_API int uf_profile_enable(uf_model* m, int on); /* on: event pair around every launch (serialises slots) */
UF_API int uf_profile_reset(uf_model* m);
UF_API int uf_profile_read(uf_model* m, uf_kernel_stat* out, uint32_t cap, uint32_t* n_out);
/* number of kernel launches issued by this handle since load (the library's own kernels only) */
UF_API int uf_launch_count(const uf_model* m, uint64_t* n);

/* fault injection for the error-path tests: every following uf_infer_batch* call on this handle fails with UF_ERR_CUDA
 * after `stages` pipeline stages have been submitted (stages < 0: off). A failed call leaves the handle usable. */
UF_API int uf_debug_fail_after(uf_model* m, int32_t stages);

/* ---- host-only entry points (no GPU needed): loader / lowering / tap tables ---- */
/* Parses + lowers the ONNX file and writes a JSON description (ops after folding with weight
 * checksums, heads, prior count, MACs, bytes) into out[cap]; *needed = bytes incl. NUL. */
UF_API int uf_onnx_inspect(const char* onnx_path, uint32_t net_w, uint32_t net_h, char* out, size_t cap,
                           size_t* needed);
/* One axis of the triangle resize (image 0.24.5 sample.rs): returns max taps via *max_taps; when
 * left/ntaps/w are non-NULL fills left[dst_len], ntaps[dst_len], w[dst_len * w_pitch]. */
UF_API int uf_resize_taps(uint32_t src_len, uint32_t dst_len, int32_t* left, int32_t* ntaps, float* w,
                          uint32_t w_pitch, uint32_t* max_taps);

/* ==== JPEG in front of the path ================================================================
 * Replaces `turbojpeg::decompress_image` (inferer.rs:35): frames arrive as baseline JPEG (what a V4L2 MJPG webcam sends,
 * cam_sender/src/sensors.rs); the host parses the headers and removes the 0xFF00 byte stuffing, the entropy-coded bytes
 * cross PCIe as they are (75 KB instead of 3 bytes per pixel) and Huffman decoding, dequantisation, inverse DCT, chroma
 * upsampling and YCbCr->RGB run on the GPU, bit-exact with libjpeg-turbo's default decoder (ISLOW IDCT, fancy upsampling);
 * the decoded frame goes straight into the resize. Frames with restart intervals, markers inside the scan, or data that
 * does not decode to exactly the frame's blocks are Huffman-decoded on host threads instead (same results; what crosses
 * PCIe is then the list of nonzero coefficients), as is everything under UF_FLAG_JPEG_HOST_HUFFMAN. Baseline /
 * extended-sequential Huffman, 8 bit, one interleaved scan, grey or YCbCr 4:4:4 / 4:2:2 / 4:2:0; anything else
 * (progressive, arithmetic, CMYK ...) is UF_ERR_UNSUPPORTED. */
typedef struct uf_jpeg_info {
    uint32_t w, h, ncomp;
    uint32_t hs[3], vs[3];      /* sampling factors */
    uint32_t nblocks;           /* 8x8 blocks in decode (MCU-interleaved) order */
    uint32_t nonzero;           /* nonzero coefficients (uf_jpeg_coefficients only) */
    uint16_t quant[3][64];      /* per component, natural (row-major) order (uf_jpeg_coefficients only) */
} uf_jpeg_info;
/* jpeg[i] / len[i]: one JPEG file per frame. Otherwise as uf_infer_batch. */
UF_API int uf_infer_batch_jpeg(uf_model* m, const uint8_t* const* jpeg, const size_t* len, uint32_t n, uf_det* out, uint32_t cap,
                               uint32_t* n_out);
/* parity hook: the quantised coefficients as the DEVICE Huffman decoder produces them (same layout as uf_jpeg_coefficients);
 * *on_device = 1 + the synchronisation rounds it took if the frame went through it, 0 if it was handed to the host decoder
 * (restart intervals, damaged data). */
UF_API int uf_jpeg_coefficients_gpu(uf_model* m, const uint8_t* jpeg, size_t len, int16_t* coefs, size_t cap_blocks, int32_t* on_device);
/* parity hook: the decoded RGB8 pixels of one frame, as the GPU kernels produce them (out_rgb: w * h * 3 bytes, cap_bytes
 * its capacity; *w / *h are set even when the capacity is too small: UF_ERR_CAPACITY). */
UF_API int uf_jpeg_decode_rgb(uf_model* m, const uint8_t* jpeg, size_t len, uint8_t* out_rgb, size_t cap_bytes, uint32_t* w,
                              uint32_t* h);
/* host only: headers. */
UF_API int uf_jpeg_info_read(const uint8_t* jpeg, size_t len, uf_jpeg_info* out);
/* host only: Huffman decoding alone. coefs (nullable) = nblocks x 64 quantised coefficients, blocks in decode order, natural
 * order inside a block; cap_blocks = its capacity in blocks. */
UF_API int uf_jpeg_coefficients(const uint8_t* jpeg, size_t len, uf_jpeg_info* info, int16_t* coefs, size_t cap_blocks);

/* ==== after the path: overlay + JPEG encode ====================================================
 * Replaces `draw_bboxes_on_image` + `turbojpeg::compress_image(&frame, 95, Subsamp::Sub2x2)` (inferer.rs:38-39, 58-92).
 * RECTANGLES: every detection's outline is drawn as imageproc's `draw_hollow_rect` draws it — corners from the reference's
 * casts (x_tl as i32, (x_br - x_tl) as u32 ... of bbox * (scale_w, scale_h); the reference passes 1280 x 720 whatever the
 * frame, router.rs:66-67), four clipped 1-pixel segments in (0, 255, 0); a box whose width or height casts to 0 is skipped
 * (the reference would panic in Rect::of_size).
 * TEXT (inferer.rs:80-88, `draw_text(.., x_tl as i32, y_tl as i32, Scale 16, DejaVuSansMono, "{:.2}%" of confidence * 100)`):
 * drawn once a glyph atlas is set (uf_text_atlas_set below), detection by detection in the reference's order (rectangle,
 * then its text), each glyph pixel blended as imageproc's draw_text_mut does: weighted_sum(pixel, colour, 1 - v, v) per
 * channel in f32, clamped, truncated to u8. RASTERISING the font stays with the reference's own rasteriser: the binding
 * renders the characters of the text once per caret position with rusttype at start-up and hands the coverage over
 * (INTEGRATION.md shows the 20 lines of Rust); placing, clipping and blending are done here. Without an atlas no text.
 * The frame is then encoded as baseline JPEG, YCbCr 4:2:0, Annex K Huffman tables, jpeg_set_quality(quality) tables: colour
 * conversion, chroma downsampling, forward ISLOW DCT and quantisation on the GPU (the coefficients are libjpeg-turbo's, bit
 * for bit), Huffman coding on the host. out / cap: destination buffer; *out_len = bytes needed (UF_ERR_CAPACITY if > cap). */
UF_API int uf_annotate_encode_jpeg(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, const uf_det* dets, uint32_t n_dets,
                                   float scale_w, float scale_h, uint32_t quality, uint8_t* out, size_t cap, size_t* out_len);
/* Same with the frame given as the baseline JPEG it arrived as (decoded on the GPU, N2; never leaves the device as pixels). */
UF_API int uf_annotate_reencode_jpeg(uf_model* m, const uint8_t* jpeg, size_t len, const uf_det* dets, uint32_t n_dets,
                                     float scale_w, float scale_h, uint32_t quality, uint8_t* out, size_t cap, size_t* out_len);
/* Batch form, the whole of inferer.rs:35-39 for n frames on the GPU: Huffman decoding, IDCT, colour (N2), rectangles + text,
 * colour conversion, downsampling, forward DCT, quantisation, Huffman CODING and byte stuffing; the host parses the incoming
 * headers and writes the outgoing ones. dets = all frames' detections back to back, det_counts[i] of them belong to frame i.
 * Frame i's file is written at out + i * out_stride, out_len[i] bytes; UF_ERR_CAPACITY if one does not fit (every out_len is
 * still set). The files are byte for byte the ones uf_annotate_reencode_jpeg writes. */
UF_API int uf_annotate_reencode_batch_jpeg(uf_model* m, const uint8_t* const* jpeg, const size_t* len, uint32_t n, const uf_det* dets,
                                           const uint32_t* det_counts, float scale_w, float scale_h, uint32_t quality, uint8_t* out,
                                           size_t out_stride, size_t* out_len);
/* The whole body of the reference's worker loop for n frames in one call (inferer.rs:35-46: decompress_image -> infer_faces ->
 * draw_bboxes_on_image -> compress_image): the frames are decoded once, the hot path runs on the decoded pixels where they
 * lie in device memory, the detections found (at most cap per frame, dets[i * cap ..], n_dets[i] = how many were selected) are
 * drawn and the annotated frames encoded. Output buffers as above. */
UF_API int uf_worker_batch_jpeg(uf_model* m, const uint8_t* const* jpeg, const size_t* len, uint32_t n, float scale_w, float scale_h,
                                uint32_t quality, uf_det* dets, uint32_t cap, uint32_t* n_dets, uint8_t* out, size_t out_stride,
                                size_t* out_len);
/* host only: the Huffman-coding / file-writing half alone. coefs = quantised blocks of a w x h YCbCr 4:2:0 frame, per
 * component plane (Y, Cb, Cr; each padded to whole 16x16 MCUs) in raster order, natural order inside a block; tables =
 * jpeg_set_quality(quality). uf_jpeg_quality_tables returns those tables (natural order). */
UF_API int uf_jpeg_write_coefficients(uint32_t w, uint32_t h, uint32_t quality, const int16_t* coefs, size_t n_blocks, uint8_t* out,
                                      size_t cap, size_t* out_len);
UF_API int uf_jpeg_quality_tables(uint32_t quality, uint16_t* lum64, uint16_t* chr64);
/* Glyph atlas for the text overlay. glyphs[pos * n_chars + k] = the glyph of charset[k] as the pos-th character of a text:
 * x0 / y0 = its pixel_bounding_box().min relative to the text origin (what rusttype's layout gives for a caret started at
 * (0, ascent)), w x h = the box, coverage[offset + gy * w + gx] = the value rusttype's draw() reports for pixel (gx, gy).
 * w == 0: nothing to draw (a blank). Characters outside the charset and positions >= max_len are skipped. n_chars = 0 removes
 * the atlas. */
typedef struct uf_glyph {
    int32_t x0, y0;
    uint32_t w, h;
    uint32_t offset;
} uf_glyph;
UF_API int uf_text_atlas_set(uf_model* m, const char* charset, uint32_t n_chars, uint32_t max_len, const uf_glyph* glyphs,
                             const float* coverage, size_t n_coverage);
/* host only: the text the overlay prints for a confidence — Rust's format!("{:.2}%", confidence * 100.0) on f32 (the product
 * rounded to f32, its exact value printed with two decimals). Writes a NUL-terminated string (cap >= 16). */
UF_API int uf_confidence_text(float confidence, char* out, size_t cap);
/* parity hook: the frame with the overlay drawn (RGB8, w * h * 3 bytes). */
UF_API int uf_draw_boxes_rgb(uf_model* m, const uint8_t* rgb, uint32_t w, uint32_t h, const uf_det* dets, uint32_t n_dets,
                             float scale_w, float scale_h, uint8_t* out_rgb);

/* ==== stream batcher + stream -> GPU routing ==========================================
 * Replaces the loop of `Inferer::run` (inferer.rs:29-50: recv_ref -> model.run -> send) and the bounded LOSSY queue in
 * front of it (`INFER_IMAGES_CHANNEL`, capacity 10, lib.rs:32-37; the router fills it with `try_send_ref` and drops the
 * frame when it is full, router.rs:64-72). One batcher owns one model handle per device; a frame of stream `s` (the
 * `hashed(&id)` of router.rs:58, see uf_stream_hash) always goes to device `devices[s % n_devices]`, so streams shard
 * over the GPUs with no cross-GPU traffic and the results of one stream stay in submission order. Per device, worker
 * threads drain up to `max_batch` frames — or what has arrived when `max_delay_us` expires, so a lone webcam is not held
 * back — call the batched path with several batches in flight (the handle's lanes), and queue the detections for
 * uf_batcher_poll, batch by batch in formation order.
 * Frames live in pinned host memory owned by the batcher, one pool per device: uf_batcher_acquire hands the ingest side a
 * slot of the right pool to decode INTO (no extra host copy; N4), uf_batcher_commit queues it. uf_batcher_try_submit is
 * acquire + memcpy + commit for callers that already hold the pixels. None of these block. */
typedef struct uf_batcher uf_batcher;

typedef struct uf_batcher_config {
    uint32_t struct_size;        /* = sizeof(uf_batcher_config) */
    uf_config model;             /* template for every device's handle (`device` is overridden; max_batch 0 = this max_batch) */
    const int32_t* devices;      /* CUDA ordinals, one handle each (an ordinal may repeat); NULL = {0} */
    uint32_t n_devices;
    uint32_t max_batch;          /* frames per batch (0 = 64) */
    uint32_t max_delay_us;       /* how long a worker waits for a batch to fill (0 = 2000) */
    uint32_t capacity;           /* frames queued per device before new ones are dropped (0 = 2 * max_batch) */
    uint32_t workers;            /* batches in flight per device (0 = 2) */
    uint32_t det_cap;            /* detections returned per frame (0 = 64) */
    uint32_t max_frame_bytes;    /* pinned slot size (0 = 1280 * 720 * 3) */
    /* the rest of the worker loop (inferer.rs:38-46): frames submitted as JPEG come back annotated and re-encoded */
    uint32_t annotate_quality;   /* 0 = detections only; 1..100 = draw the detections and encode at this quality (the reference: 95) */
    float annotate_scale_w;      /* what the relative boxes are multiplied by (0 = 1280, router.rs:66) */
    float annotate_scale_h;      /* (0 = 720, router.rs:67) */
    uint32_t annotate_max_bytes; /* room per annotated file (0 = 1 MiB); a larger file fails its frame with UF_ERR_CAPACITY */
} uf_batcher_config;

typedef struct uf_result {
    uint64_t stream;             /* as submitted */
    uint64_t user_tag;           /* as submitted (e.g. a frame sequence number) */
    int32_t device;              /* CUDA ordinal that ran the frame */
    int32_t status;              /* UF_OK, or the status of the failed batch (the frame is skipped, inferer.rs:37) */
    uint32_t n_dets;             /* faces selected (may exceed det_cap; only min(n_dets, det_cap) were returned) */
    uint32_t batch_size;         /* frames in the batch this frame rode in */
    uint64_t latency_us;         /* commit -> result queued */
    uint32_t file_bytes;         /* annotate mode: size of the frame's annotated JPEG (uf_batcher_poll_frames), else 0 */
    uint32_t reserved_;
} uf_result;

typedef struct uf_batcher_stats {
    uint64_t submitted, dropped, completed, failed, batches;
} uf_batcher_stats;

UF_API int uf_batcher_create(const uf_batcher_config* cfg, uf_batcher** out);
/* Same with the batched call injected (the seam `trait InferModel` is in the reference, nn.rs:24-26): `fn` receives what
 * uf_infer_batch would, plus `user` and the device ordinal; no handle is loaded and CUDA is not touched. This is how the
 * queueing / routing / ordering logic is tested without a GPU; a product batcher passes through uf_batcher_create. */
typedef int (*uf_batch_fn)(void* user, int32_t device, const uint8_t* const* rgb, const uint32_t* w, const uint32_t* h,
                           uint32_t n, uf_det* out, uint32_t cap, uint32_t* n_out);
UF_API int uf_batcher_create_ex(const uf_batcher_config* cfg, uf_batch_fn fn, void* user, uf_batcher** out);
UF_API void uf_batcher_destroy(uf_batcher* b); /* finishes what is queued, joins the workers, frees the handles */
/* *buf = pinned slot of >= bytes in the owner device's pool, or NULL when the queue is full (frame dropped, counted). */
UF_API int uf_batcher_acquire(uf_batcher* b, uint64_t stream, size_t bytes, uint8_t** buf, uint64_t* ticket);
UF_API int uf_batcher_commit(uf_batcher* b, uint64_t ticket, uint32_t w, uint32_t h, uint64_t user_tag);
UF_API int uf_batcher_abort(uf_batcher* b, uint64_t ticket);
UF_API int uf_batcher_try_submit(uf_batcher* b, uint64_t stream, const uint8_t* rgb, uint32_t w, uint32_t h,
                                 uint64_t user_tag, int32_t* accepted);
/* The same for frames that arrive as JPEG files (N2; what `FrameMsg.data` holds, common/src/protocol.rs:15-19): the slot
 * receives the file's bytes, the worker sends the batch through uf_infer_batch_jpeg. RGB and JPEG frames may be mixed. */
UF_API int uf_batcher_commit_jpeg(uf_batcher* b, uint64_t ticket, size_t jpeg_len, uint64_t user_tag);
UF_API int uf_batcher_try_submit_jpeg(uf_batcher* b, uint64_t stream, const uint8_t* jpeg, size_t len, uint64_t user_tag,
                                      int32_t* accepted);
/* The whole ingest step for one length-delimited data-socket frame (data_socket.rs:34-47 -> router.rs:56-72): parse the
 * bincode ProtoMsg, key the stream with hashed(&id), and — for a FrameMsg — queue its JPEG payload on the GPU that owns the
 * stream. *stream (nullable) receives the key; a ConnectReq is accepted = 0 with UF_OK. */
UF_API int uf_batcher_ingest(uf_batcher* b, const uint8_t* msg, size_t len, uint64_t user_tag, int32_t* accepted, uint64_t* stream);
/* Up to cap finished frames: res[i] + dets[i * det_cap ..]. Waits at most timeout_ms for the first one. */
UF_API int uf_batcher_poll(uf_batcher* b, uf_result* res, uf_det* dets, uint32_t cap, uint32_t timeout_ms, uint32_t* n_out);
/* annotate mode: the same, and the annotated JPEG of frame i at files + i * file_stride (res[i].file_bytes bytes);
 * file_stride >= the configuration's annotate_max_bytes. Glyph atlas: uf_text_atlas_set on uf_batcher_model(b, slot). */
UF_API int uf_batcher_poll_frames(uf_batcher* b, uf_result* res, uf_det* dets, uint8_t* files, size_t file_stride, uint32_t cap,
                                  uint32_t timeout_ms, uint32_t* n_out);
UF_API int uf_batcher_flush(uf_batcher* b, uint32_t timeout_ms); /* returns when everything committed so far is pollable */
UF_API int uf_batcher_stats_read(const uf_batcher* b, uf_batcher_stats* out);
UF_API int uf_batcher_owner(const uf_batcher* b, uint64_t stream, int32_t* device);
UF_API int uf_batcher_model(uf_batcher* b, uint32_t device_slot, uf_model** out); /* borrowed: parity hooks, profiling */

/* measurement aid: `producers` C++ threads submit `total` frames (frame i = frames[i % n_frames], stream = streams[i % n_streams];
 * dropped frames are retried so that every frame is counted) while the caller's thread polls; *seconds = wall time. */
UF_API int uf_debug_batcher_drive(uf_batcher* b, const uint8_t* frames, uint32_t n_frames, uint32_t w, uint32_t h,
                                  const uint64_t* streams, uint32_t n_streams, uint64_t total, uint32_t producers, double* seconds,
                                  uint64_t* detections);

/* The same for wire messages: message i = msgs[i % n_msgs] (bincode ProtoMsg::FrameMsg carrying a JPEG), handed to
 * uf_batcher_ingest — parse, key the stream, queue the payload on its GPU; decode + detection in the batcher's workers. */
UF_API int uf_debug_batcher_drive_msgs(uf_batcher* b, const uint8_t* const* msgs, const size_t* lens, uint32_t n_msgs, uint64_t total,
                                       uint32_t producers, double* seconds, uint64_t* detections);

/* ---- ingest helpers, host only (N4) ----
 * `hashed(&id)` of infer_server/src/lib.rs:39-46: Rust's DefaultHasher (SipHash-1-3, zero keys) over the bytes of the
 * stream name followed by the 0xff terminator `str::hash` appends. */
UF_API int uf_stream_hash(const uint8_t* name, size_t len, uint64_t* out);
/* One length-delimited frame of the data socket (data_socket.rs:34-47) = a bincode-1.3 `ProtoMsg`
 * (common/src/protocol.rs:7-19): u32 LE variant (0 ConnectReq(String), 1 FrameMsg{id: String, data: Vec<u8>}),
 * u64 LE lengths. Zero-copy: *id / *data point into msg. kind: 0 / 1. Malformed input = UF_ERR_INVALID_ARG. */
UF_API int uf_protomsg_parse(const uint8_t* msg, size_t len, uint32_t* kind, const uint8_t** id, size_t* id_len,
                             const uint8_t** data, size_t* data_len);
/* test hook: SipHash-c-d with explicit keys (uf_stream_hash is c = 1, d = 3, zero keys; c = 2, d = 4 has published vectors) */
UF_API int uf_debug_siphash(uint32_t c, uint32_t d, uint64_t k0, uint64_t k1, const uint8_t* in, size_t len, uint64_t* out);

UF_API const char* uf_last_error(void);
UF_API const char* uf_version(void);
UF_API int uf_device_count(int32_t* n);

#ifdef __cplusplus
}
#endif
#endif /* ULTRAFACE_B200_H */
