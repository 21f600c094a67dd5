// kernels.h — launchers of the sm_100a kernels (K1..K11 of SURVEY.md §2).
// Activations are NHWC fp32; the graph input is HWC u8 (RgbImage layout, nn.rs:24-26).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace uf {

// NHWC fp32 view: addr(n,y,x,c) = p + n*frame_stride + (y*W + x)*pix_stride + c
struct TView {
    float* p = nullptr;
    long long frame_stride = 0;
    int pix_stride = 0;
    int C = 0, H = 0, W = 0;
};

// HWC u8 frames (3 channels, dense)
struct U8View {
    const uint8_t* p = nullptr;
    long long frame_stride = 0;
    int H = 0, W = 0;
};

struct ConvParams {
    int cin, cout, k, stride, pad, dil, groups, relu;
};

// Device-side tap tables of one (src,dst) size pair — image 0.24.5 sample.rs, see resize_taps.h
struct ResizeTapsDev {
    const int* vleft; const int* vn; const float* vw; int vmax;
    const int* hleft; const int* hn; const float* hw; int hmax;
    int tile_w, tile_h, max_cols;  // CTA tile of the destination and widest source span of a tile
    int small_w, small_cols;       // narrower tile for launches too small to fill the GPU with wide ones (0: none)
};

// ---- K1: bit-exact triangle resize (nn.rs:74-80), frames HWC u8 -> HWC u8
void launch_resize(const uint8_t* src, long long src_frame_stride, int sw, int sh, uint8_t* dst,
                   long long dst_frame_stride, int dw, int dh, int frames, const ResizeTapsDev& t,
                   int round_intermediate, cudaStream_t s);
// ---- K2 (parity hook): LUT normalise + HWC->NCHW (nn.rs:82-91)
void launch_normalise_nchw(const uint8_t* hwc, int w, int h, const float* lut, float* out, cudaStream_t s);

// ---- convolutions. `lut` (3x256 floats) is used when the input is the u8 frame.
void launch_conv_generic(const TView& in, const U8View* in_u8, const float* lut, const TView& out,
                         const TView* res, const float* w_kkio, const float* b, const ConvParams& p,
                         int frames, cudaStream_t s);
// K3 stem: 3x3 s2 p1, 3 -> 16, u8 input through the LUT, bias + ReLU
// host_w: 27*16 weights [ky][kx][ci][co] + 16 biases in HOST memory (passed as a __grid_constant__ parameter)
void launch_stem(const U8View& in, const float* lut, const TView& out, const float* host_w, int relu, int frames,
                 cudaStream_t s);
// K1 + K2 + K3 fused for frames at exactly twice the network size (kernels_prestem.cu): src = the u8 frames themselves
bool prestem_supported(const uint8_t* src, long long src_frame_stride, int sw, int sh, int net_w, int net_h, const ResizeTapsDev& t);
void launch_prestem(const U8View& src, const float* lut, const ResizeTapsDev& t, const TView& out, const float* host_w, int relu,
                    int round_intermediate, int frames, uint8_t* dbg_resized, cudaStream_t s);
// K4 depthwise 3x3 (pad 1, stride 1|2), weights [9][C]
void launch_depthwise(const TView& in, const TView& out, const float* w_tc, const float* b, int stride,
                      int relu, int frames, cudaStream_t s);
// K5 pointwise 1x1, weights [Cin][Cout]; optional residual added before ReLU
void launch_pointwise(const TView& in, const TView& out, const TView* res, const float* w_io,
                      const float* b, int relu, int frames, cudaStream_t s);
// K4+K5 fused: depthwise 3x3 (+bias+ReLU) kept in shared memory -> pointwise 1x1 (+bias, +-ReLU)
bool fused_dwpw_supported(int C, int N);
void launch_fused_dwpw(const TView& in, const TView& out, const float* dw_w_tc, const float* dw_b,
                       int stride, int dw_relu, const float* pw_w_io, const float* pw_b, int pw_relu, int frames,
                       cudaStream_t s);
// pixel-per-thread form of the fused kernel for C in {16,32,64} (large, memory-bound maps)
bool fused_dwpw_pix_supported(int C, int N, int stride);
void launch_fused_dwpw_pix(const TView& in, const TView& out, const float* dw_w_tc, const float* dw_b,
                           int stride, int dw_relu, const float* pw_w_io, const float* pw_b, int pw_relu,
                           int frames, cudaStream_t s);
// SSD heads on the 64-channel map: dw3x3 -> 1x1 (N <= 16) with all weights as kernel-parameter constants
bool head_dwpw_supported(int C, int N, int stride);
size_t head_dwpw_weight_floats(int C, int N);
// two heads (<= 8 and <= 16 outputs) of the same input in one launch, C = 64 / 128
bool head2_dwpw_supported(int C, int Na, int Nb);
void launch_head2_dwpw(const TView& in, const TView& out_a, const TView& out_b, const float* host_w, int relu_bits, int frames,
                       cudaStream_t s);
void launch_head_dwpw(const TView& in, const TView& out, const float* host_w, int dw_relu, int pw_relu, int frames,
                      cudaStream_t s);
// K6b warp-per-pixel 3x3 (stride 1, pad = dil) for many input channels and <= 16 outputs (last SSD heads)
bool conv3x3_warp_supported(int cin, int cout);
void launch_conv3x3_warp(const TView& in, const TView& out, const float* w_kkio, const float* b, int dil,
                         int relu, int frames, cudaStream_t s);
// K6 on tcgen05 (kernels_tc.cu): dense 3x3, stride 1, pad = dil, Cin / Cout <= 16; zero-copy im2col from one staged tile
bool dense3x3_tc_supported(int cin, int cout, int dil);
size_t dense3x3_tc_weight_floats(int cin);
void launch_dense3x3_tc(const TView& in, const TView& out, const float* w_hi, const float* w_lo, const float* host_bias, int dil,
                        int relu, int frames, cudaStream_t s);
// K6 small dense 3x3 (stride 1, pad = dil), Cin,Cout in {8,12,16}, weights [3][3][Cin][Cout]
bool small_dense_supported(int cin, int cout);
// host_w: 9*Cin*Cout weights [ky][kx][ci][co] + Cout biases in HOST memory (kernel-parameter constants); dil <= 8
void launch_small_dense(const TView& in, const TView& out, const float* host_w, int dil, int relu, int frames,
                        cudaStream_t s);
// K5 on tcgen05 (3xTF32, fp32-level accuracy): TMA-fed, TMEM accumulators, warp-specialised (kernels_tc.cu)
struct alignas(64) TmaMap { unsigned char bytes[128]; };  // mirrors CUtensorMap
bool make_tmap_f32_2d(TmaMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows);
bool pointwise_tc_supported(int K, int N);
int pointwise_tc_n_umma(int N);
// tm_a: activations [frames*H*W][K] (row pitch pix_stride), box 128 rows; tm_whi/tm_wlo: weights [N][K], box n_umma rows
bool make_tmap_f32_2d_store(TmaMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes);
// tm_out (nullable): output [frames*H*W][N] for the TMA-store epilogue; host_bias: N floats in HOST memory
// Depthwise mode: the A operand of the GEMM is computed on the fly, A = relu(dw3x3(in, stride) + b): `in` is the
// depthwise INPUT, `w` [9][C] / `b` [C] its DEVICE weights; the depthwise result never reaches memory.
struct TcDepthwise {
    TView in;
    const float* w;
    const float* b;
    int stride, relu;
};
// tm_res (nullable): residual tensor, same geometry as tm_out (TMA-loaded in the epilogue); dw (nullable): depthwise
// mode (tm_a is then unused and `in` only provides K)
void launch_pointwise_tc(const TmaMap& tm_a, const TmaMap& tm_whi, const TmaMap& tm_wlo, const TmaMap* tm_out,
                         const TmaMap* tm_res, const TView& in, const TView& out, const TView* res,
                         const float* host_bias, int relu, int frames, cudaStream_t s,
                         const TcDepthwise* dw = nullptr);
// K4+K5 fused, TMA-pipelined persistent form (C in {16,32,64}, N in {32,64}); tm_in: 4-D map of the input view
// with box (16, in_w, in_h, 1) and 64B swizzle; tm_out: 4-D map of the output view with box (32, out_w, out_h, 1), 128B swizzle
bool make_tmap_nhwc(TmaMap* out, const TView& v, int frames, uint32_t box_c, uint32_t box_w, uint32_t box_h, int swizzle_bytes);
bool fused_dwpw_tma_supported(int C, int N, int stride);
bool fused_dwpw_tc_supported(int C, int N, int stride);
void fused_dwpw_tma_boxes(int stride, int* in_w, int* in_h, int* out_w, int* out_h);
size_t fused_dwpw_tma_weight_floats(int C, int N);
// tensor-core form (1x1 conv on tcgen05, 3xTF32): CTA tile 8 x 16, weights [N][C] as tf32 hi / lo tensor maps
void fused_dwpw_tc_boxes(int stride, int* in_w, int* in_h, int* out_w, int* out_h);
int fused_dwpw_tc_slice_channels(int C);  // channels per input TMA box (its swizzle span is 4x that in bytes)
bool make_tmap_f32_2d_sw(TmaMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                         uint32_t box_cols, uint32_t box_rows);
void launch_fused_dwpw_tc(const TmaMap& tm_in, const TmaMap& tm_out, const TmaMap& tm_whi, const TmaMap& tm_wlo, const TView& in,
                          const TView& out, const float* host_w, int stride, int dw_relu, int pw_relu, int frames,
                          cudaStream_t s);
// host_w: [dw 9*C][dw bias C][pw C*N][pw bias N] in HOST memory (passed by value as a __grid_constant__ parameter)
void launch_fused_dwpw_tma(const TmaMap& tm_in, const TmaMap& tm_out, const TView& in, const TView& out,
                           const float* host_w, int stride, int dw_relu, int pw_relu, int frames, cudaStream_t s);
// fallbacks for graphs that do not fuse completely
void launch_add(const TView& a, const TView& b, const TView& out, int relu, int frames, cudaStream_t s);
void launch_relu(const TView& a, const TView& out, int frames, cudaStream_t s);
void launch_copy(const TView& a, const TView& out, int frames, cudaStream_t s);
// NHWC view -> dense NCHW (uf_tensor_read)
void launch_nhwc_to_nchw(const TView& a, int frame, float* out, cudaStream_t s);

// programmatic dependent launch for the kernel chain (pdl.cuh); process-wide switch, default off
bool pdl_enabled();
void pdl_set_enabled(bool on);

// ---- K8: softmax over 2 classes + prior-box decode; conf [F,K,2], loc [F,K,4] raw head outputs
void launch_tail(const float* conf, const float* loc, long long conf_frame_stride,
                 long long loc_frame_stride, const float* priors, int K, float center_var,
                 float size_var, float* scores, float* boxes, int frames, cudaStream_t s);

// ---- N2: JPEG sample-domain decode (kernels_jpeg.cu): planes = scratch for the component planes, rgb = destination
struct JpegPlan;
struct JpegBatchDev {
    const JpegPlan* plans;     // [frames]
    const uint32_t* offs;      // block offsets of all frames (JpegPlan::offs_base)
    const uint32_t* entries;   // (natural index << 16 | value) of all frames (JpegPlan::entries_base)
    uint8_t* planes;           // JpegPlan::planes_off
    uint8_t* rgb;              // JpegPlan::rgb_off
    const int16_t* dcv;        // non-NULL (GPU Huffman path): DC values [JpegPlan::offs_base + block]; the lists then hold the AC part only
    const int* status;         // non-NULL (GPU Huffman path): frames with a nonzero status are skipped (their lists are not valid)
};
// Huffman decoding on the GPU (kernels_jpeg_huff.cu)
struct JpegHuffFrame;
struct JpegHuffTabSet;
struct JpegHuffBatch {
    const JpegHuffFrame* frames;        // [frames]
    const JpegHuffTabSet* tabsets;      // the batch's distinct table sets (JpegHuffFrame::tabset)
    const uint8_t* bytes;               // unstuffed entropy-coded segments (JpegHuffFrame::data_off)
    unsigned long long* start_used;     // [subsequences] start state of the last decode of each subsequence
    uint32_t* counts;                   // [subsequences] blocks started in it | nonzero AC coefficients in it << 16
    uint32_t* offs;                     // [blocks + frames] first AC entry of each block, per frame nblocks + 1 (JpegHuffFrame::offs_base)
    uint32_t* entries;                  // AC entries of all frames (JpegHuffFrame::ent_base)
    int16_t* dcv;                       // [blocks + frames] DC differences, then DC values
    int* status;                        // [frames] zeroed by the caller; stays 0 = settled and decoded to exactly its blocks
};
void launch_jhuff_sync(const JpegHuffBatch& b, int frames, uint32_t max_nsub, int first, const unsigned long long* in,
                       unsigned long long* out, cudaStream_t s);
int jhuff_rounds(uint32_t max_nsub);  // launches of launch_jhuff_sync that make every frame of the run exact
void launch_jhuff_finish(const JpegHuffBatch& b, int frames, uint32_t max_nsub, const unsigned long long* fin, cudaStream_t s);
void launch_jpeg_decode(const JpegBatchDev& b, int frames, uint32_t max_nblocks, uint32_t max_w, uint32_t max_h, cudaStream_t s);

// ---- N3: rectangle overlay + JPEG encode front half (kernels_jpeg_enc.cu)
void launch_draw_rects(uint8_t* rgb, int w, int h, const int4* d_rects, int n, cudaStream_t s);
struct OverlayGlyph {  // one placed glyph: top-left pixel in the image, box, offset of its coverage values in the atlas
    int32_t x, y;
    uint32_t w, h, offset;
};
struct OverlayFrame {  // batch form: one frame's share of the concatenated lists
    unsigned long long rgb_off;  // byte offset of the frame's pixels
    int32_t w, h;
    uint32_t rect_first, n;      // rectangles (= detections drawn)
    uint32_t gstart_first, glyph_first;
    uint32_t text;               // 0: rectangles only
    uint32_t pad_;
};
void launch_draw_overlay_batch(uint8_t* rgb_base, const OverlayFrame* d_frames, int frames, const int4* d_rects, const uint32_t* d_glyph_start,
                               const OverlayGlyph* d_glyphs, const float* d_coverage, cudaStream_t s);
void launch_draw_overlay(uint8_t* rgb, int w, int h, const int4* d_rects, const uint32_t* d_glyph_start, const OverlayGlyph* d_glyphs,
                         const float* d_coverage, int n, cudaStream_t s);
void launch_jpeg_encode(const uint8_t* d_rgb, const JpegPlan& plan, uint8_t* d_planes, int16_t* d_coefs, cudaStream_t s);
struct JpegEncJob;
void launch_jpeg_encode_batch(const uint8_t* d_rgb_base, const JpegEncJob* d_jobs, int frames, uint32_t max_cw, uint32_t max_ch, uint32_t max_blocks,
                              uint8_t* d_planes_base, int16_t* d_coefs_base, cudaStream_t s);

// ---- N3: Huffman coding of quantised 4:2:0 frames on the GPU (kernels_jpeg_henc.cu)
struct JpegEncTables;      // jpeg_decode.h
struct JpegEncFrame;       // jpeg_decode.h
struct JpegEncBatch {
    const JpegEncFrame* frames;
    const int16_t* coefs;
    const JpegEncTables* tables;
    uint32_t* bitlen;      // [blocks of all frames]
    uint32_t* bitoff;
    uint32_t* frame_bits;  // [frames]
    uint32_t* packed;      // zeroed by the caller
    uint8_t* out;
    uint32_t* out_len;     // [frames] bytes of the entropy-coded segment; 0xffffffff: did not fit (the host encoder takes the frame)
};
void launch_jpeg_huffman_encode(const JpegEncBatch& B, int frames, uint32_t max_nblocks, cudaStream_t s);

// ---- K9-K11: threshold + sort + greedy NMS, one CTA per frame (nn.rs:109-140,198-260)
struct PostBuffers {
    unsigned long long* sort_scratch;  // [frames][sort_cap] keys, used when candidates exceed smem
    int sort_cap;                      // power of two >= K
    float* sel_boxes;                  // [frames][K][4] selected boxes (16-byte aligned working copy)
    float* dets;                       // [frames][K][5]
    int* det_idx;                      // [frames][K] prior index of each detection (nullable)
    int* counts;                       // [frames]
    // bit-matrix NMS for frames with many candidates (nullable: then every frame is resolved inside post_kernel)
    unsigned* mask;                    // [frames][K][mask_pitch] suppression bits, rows / columns in processing order
    int mask_pitch;                    // words per row (post_mask_pitch(K))
    int* big_n;                        // [frames] candidates of a frame left to the bit-matrix kernels (0: done already)
    int* any_big;                      // [1] some frame of the stage published a candidate list (reset by the sweep)
};
void launch_post(const float* scores, const float* boxes, int K, float min_conf, float max_iou,
                 const PostBuffers& pb, int frames, cudaStream_t s);
void launch_tail_post(const float* conf, const float* loc, long long conf_frame_stride, long long loc_frame_stride,
                      const float* priors, float center_var, float size_var, float* scores, float* boxes, int K,
                      float min_conf, float max_iou, const PostBuffers& pb, int frames, cudaStream_t s);
size_t post_sort_scratch_elems(int K);  // sort_cap for a given K
int post_configure();                   // opt in to large dynamic smem; returns cudaError_t
// second half of the post step for the frames post_kernel left to the bit-matrix path (requires pb.mask)
void launch_nms_mask(int K, float max_iou, const PostBuffers& pb, int frames, cudaStream_t s);
void launch_nms_sweep(const float* scores, int K, const PostBuffers& pb, int frames, cudaStream_t s);
int post_mask_pitch(int K);
size_t post_mask_words(int K);  // per frame
bool post_mask_supported(int K);

}  // namespace uf
