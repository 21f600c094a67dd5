// kernels_jpeg_enc.cu — N3: the sample-domain half of what follows the hot path in the reference's worker,
// `draw_bboxes_on_image` + `turbojpeg::compress_image(&frame, 95, Sub2x2)` (/root/reference/infer_server/src/inferer.rs:38-39,
// 58-92), on the GPU:
//   draw_rects_kernel       imageproc `draw_hollow_rect`: four clipped 1-pixel line segments per box, colour (0, 255, 0).
//                           (The reference also prints the confidence with rusttype's anti-aliased glyphs: not drawn here.)
//   jpeg_enc_color_kernel   libjpeg-turbo `rgb_ycc_convert` (jccolor.c fixed-point tables) + `h2v2_downsample` (jcsample.c:
//                           (a+b+c+d+bias)>>2, bias alternating 1,2) + edge expansion (pixels replicated to the right, the
//                           last real component row replicated downwards) -> Y / Cb / Cr planes padded to whole MCUs
//   jpeg_fdct_kernel        `jpeg_fdct_islow` (jfdctint.c) on sample - 128 + quantisation ((|v| + q/2) / q with q = 8 * table
//                           entry, sign restored) -> int16 blocks, natural order, per plane in raster order
// Huffman coding and the file format are on the host (jpeg_encode.cc). Exact integer arithmetic: the coefficients equal
// the ones libjpeg-turbo writes for the same pixels (tests/test_jpeg_encode.py compares with PIL's files).
#include "jpeg_decode.h"
#include "kernels.h"

namespace uf {

__global__ void draw_rects_kernel(uint8_t* __restrict__ rgb, int w, int h, const int4* __restrict__ rects, int n) {
    const int4 r = rects[blockIdx.x];  // left, top, right, bottom (inclusive)
    const int rw = r.z - r.x + 1, rh = r.w - r.y + 1;
    auto put = [&](int x, int y) {
        if (x >= 0 && x < w && y >= 0 && y < h) {
            uint8_t* p = rgb + ((size_t)y * w + x) * 3;
            p[0] = 0; p[1] = 255; p[2] = 0;
        }
    };
    for (int i = threadIdx.x; i < rw; i += blockDim.x) { put(r.x + i, r.y); put(r.x + i, r.w); }
    for (int i = threadIdx.x; i < rh; i += blockDim.x) { put(r.x, r.y + i); put(r.z, r.y + i); }
}

void launch_draw_rects(uint8_t* rgb, int w, int h, const int4* d_rects, int n, cudaStream_t s) {
    if (n > 0) draw_rects_kernel<<<n, 128, 0, s>>>(rgb, w, h, d_rects, n);
}

// imageproc's Clamp<f32> for u8 after weighted_sum: below 255 and above 0 the float is truncated
__device__ __forceinline__ uint8_t blend_u8(uint8_t pix, float col, float lw, float rw) {
    const float x = __fadd_rn(__fmul_rn((float)pix, lw), __fmul_rn(col, rw));  // left * left_weight + right * right_weight, no fma
    return x < 255.0f ? (x > 0.0f ? (uint8_t)x : (uint8_t)0) : (uint8_t)255;
}

// Rectangles AND text, detection by detection in the reference's order (a later detection's rectangle overwrites an earlier
// one's text where they overlap, a glyph blends over whatever is there): one CTA walks the list, its threads share the
// pixels of the item in hand. A frame has a handful of detections; this is nowhere near the encoder's cost.
__device__ __forceinline__ void draw_overlay_frame(uint8_t* __restrict__ rgb, int w, int h, const int4* __restrict__ rects,
                                                   const uint32_t* __restrict__ glyph_start, const OverlayGlyph* __restrict__ glyphs,
                                                   const float* __restrict__ coverage, int n) {
    for (int d = 0; d < n; ++d) {
        const int4 r = rects[d];
        const int rw = r.z - r.x + 1, rh = r.w - r.y + 1;
        auto put = [&](int x, int y) {
            if (x >= 0 && x < w && y >= 0 && y < h) {
                uint8_t* p = rgb + ((size_t)y * w + x) * 3;
                p[0] = 0; p[1] = 255; p[2] = 0;
            }
        };
        for (int i = threadIdx.x; i < rw; i += blockDim.x) { put(r.x + i, r.y); put(r.x + i, r.w); }
        for (int i = threadIdx.x; i < rh; i += blockDim.x) { put(r.x, r.y + i); put(r.z, r.y + i); }
        __syncthreads();
        if (!glyph_start) continue;  // rectangles only
        for (uint32_t g = glyph_start[d]; g < glyph_start[d + 1]; ++g) {
            const OverlayGlyph G = glyphs[g];
            const uint32_t npx = G.w * G.h;
            for (uint32_t i = threadIdx.x; i < npx; i += blockDim.x) {
                const uint32_t gy = i / G.w, gx = i - gy * G.w;
                const int x = G.x + (int)gx, y = G.y + (int)gy;
                if (x < 0 || x >= w || y < 0 || y >= h) continue;
                const float v = coverage[G.offset + i], lw = __fsub_rn(1.0f, v);
                uint8_t* p = rgb + ((size_t)y * w + x) * 3;
                p[0] = blend_u8(p[0], 0.0f, lw, v);
                p[1] = blend_u8(p[1], 255.0f, lw, v);
                p[2] = blend_u8(p[2], 0.0f, lw, v);
            }
            __syncthreads();  // the next glyph's box may overlap this one's
        }
    }
}

__global__ void __launch_bounds__(256)
draw_overlay_kernel(uint8_t* __restrict__ rgb, int w, int h, const int4* __restrict__ rects, const uint32_t* __restrict__ glyph_start,
                    const OverlayGlyph* __restrict__ glyphs, const float* __restrict__ coverage, int n) {
    draw_overlay_frame(rgb, w, h, rects, glyph_start, glyphs, coverage, n);
}

// batch form: CTA = frame
__global__ void __launch_bounds__(256)
draw_overlay_batch_kernel(uint8_t* __restrict__ rgb_base, const OverlayFrame* __restrict__ frames, const int4* __restrict__ rects,
                          const uint32_t* __restrict__ glyph_start, const OverlayGlyph* __restrict__ glyphs, const float* __restrict__ coverage) {
    const OverlayFrame F = frames[blockIdx.x];
    if (F.n == 0) return;
    draw_overlay_frame(rgb_base + F.rgb_off, F.w, F.h, rects + F.rect_first, F.text ? glyph_start + F.gstart_first : nullptr,
                       glyphs + F.glyph_first, coverage, (int)F.n);
}

void launch_draw_overlay(uint8_t* rgb, int w, int h, const int4* d_rects, const uint32_t* d_glyph_start, const OverlayGlyph* d_glyphs,
                         const float* d_coverage, int n, cudaStream_t s) {
    if (n > 0) draw_overlay_kernel<<<1, 256, 0, s>>>(rgb, w, h, d_rects, d_glyph_start, d_glyphs, d_coverage, n);
}

void launch_draw_overlay_batch(uint8_t* rgb_base, const OverlayFrame* d_frames, int frames, const int4* d_rects, const uint32_t* d_glyph_start,
                               const OverlayGlyph* d_glyphs, const float* d_coverage, cudaStream_t s) {
    if (frames > 0) draw_overlay_batch_kernel<<<frames, 256, 0, s>>>(rgb_base, d_frames, d_rects, d_glyph_start, d_glyphs, d_coverage);
}

__device__ __forceinline__ void jrgb2ycc(const uint8_t* __restrict__ p, int& y, int& cb, int& cr) {
    const int r = p[0], g = p[1], b = p[2];
    y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
}

// thread = one chroma sample = a 2x2 quad of luma samples, over the PADDED planes
__device__ __forceinline__ void jpeg_enc_color_body(const uint8_t* __restrict__ rgb, const JpegPlan& plan, uint8_t* __restrict__ planes) {
    const int cw = (int)plan.plane_w[1], ch = (int)plan.plane_h[1];
    const int cx = blockIdx.x * 32 + (threadIdx.x & 31), cy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (cx >= cw || cy >= ch) return;
    const int w = (int)plan.w, h = (int)plan.h;
    // luma: every padded position takes the nearest real pixel
    int cbq[4], crq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int x = min(2 * cx + (q & 1), w - 1), y = min(2 * cy + (q >> 1), h - 1);
        int yy, cb, cr;
        jrgb2ycc(rgb + ((size_t)y * w + x) * 3, yy, cb, cr);
        planes[plan.plane_off[0] + (size_t)(2 * cy + (q >> 1)) * plan.plane_w[0] + 2 * cx + (q & 1)] = (uint8_t)yy;
    }
    // chroma: pixel columns are replicated to the right BEFORE averaging, whole chroma rows are replicated downwards AFTER it
    const int cyc = min(cy, (int)plan.real_h[1] - 1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int x = min(2 * cx + (q & 1), w - 1), y = min(2 * cyc + (q >> 1), h - 1);
        int yy;
        jrgb2ycc(rgb + ((size_t)y * w + x) * 3, yy, cbq[q], crq[q]);
    }
    const int bias = 1 + (cx & 1);
    planes[plan.plane_off[1] + (size_t)cy * cw + cx] = (uint8_t)((cbq[0] + cbq[1] + cbq[2] + cbq[3] + bias) >> 2);
    planes[plan.plane_off[2] + (size_t)cy * cw + cx] = (uint8_t)((crq[0] + crq[1] + crq[2] + crq[3] + bias) >> 2);
}

#define EFIX_0_298631336 2446
#define EFIX_0_390180644 3196
#define EFIX_0_541196100 4433
#define EFIX_0_765366865 6270
#define EFIX_0_899976223 7373
#define EFIX_1_175875602 9633
#define EFIX_1_501321110 12299
#define EFIX_1_847759065 15137
#define EFIX_1_961570560 16069
#define EFIX_2_053119869 16819
#define EFIX_2_562915447 20995
#define EFIX_3_072711026 25172

__device__ __forceinline__ int edescale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of jpeg_fdct_islow; PASS2 selects the second pass' scaling
template <bool PASS2>
__device__ __forceinline__ void jfdct_1d(const int d[8], int o[8]) {
    const int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
    const int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    constexpr int SH = PASS2 ? 13 + 2 : 13 - 2;
    if (PASS2) {
        o[0] = edescale(tmp10 + tmp11, 2);
        o[4] = edescale(tmp10 - tmp11, 2);
    } else {
        o[0] = (tmp10 + tmp11) << 2;
        o[4] = (tmp10 - tmp11) << 2;
    }
    int z1 = (tmp12 + tmp13) * EFIX_0_541196100;
    o[2] = edescale(z1 + tmp13 * EFIX_0_765366865, SH);
    o[6] = edescale(z1 + tmp12 * (-EFIX_1_847759065), SH);
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    const int z5 = (z3 + z4) * EFIX_1_175875602;
    const int t4 = tmp4 * EFIX_0_298631336, t5 = tmp5 * EFIX_2_053119869, t6 = tmp6 * EFIX_3_072711026, t7 = tmp7 * EFIX_1_501321110;
    z1 *= -EFIX_0_899976223; z2 *= -EFIX_2_562915447; z3 *= -EFIX_1_961570560; z4 *= -EFIX_0_390180644;
    z3 += z5; z4 += z5;
    o[7] = edescale(t4 + z1 + z3, SH);
    o[5] = edescale(t5 + z2 + z4, SH);
    o[3] = edescale(t6 + z2 + z3, SH);
    o[1] = edescale(t7 + z1 + z4, SH);
}

constexpr int EB_PER_CTA = 32, EB_STRIDE = 72;

__global__ void __launch_bounds__(256)
jpeg_enc_color_kernel(const uint8_t* __restrict__ rgb, JpegPlan plan, uint8_t* __restrict__ planes) {
    jpeg_enc_color_body(rgb, plan, planes);
}

// batch form: blockIdx.z = frame, the grid covers the largest frame
__global__ void __launch_bounds__(256)
jpeg_enc_color_batch_kernel(const uint8_t* __restrict__ rgb_base, const JpegEncJob* __restrict__ jobs, uint8_t* __restrict__ planes_base) {
    const JpegEncJob& J = jobs[blockIdx.z];
    jpeg_enc_color_body(rgb_base + J.rgb_off, J.plan, planes_base + J.planes_off);
}

__device__ __forceinline__ void jpeg_fdct_body(const uint8_t* __restrict__ planes, const JpegPlan& plan, int16_t* __restrict__ coefs, int* ws) {
    const int tid = threadIdx.x, t = tid & 7;
    const uint32_t blk = blockIdx.x * EB_PER_CTA + (tid >> 3);  // block index over the three planes, raster order inside each
    uint32_t nb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) nb[c] = plan.plane_w[c] * plan.plane_h[c] / 64;
    if (blk >= nb[0] + nb[1] + nb[2]) return;
    const int c = blk < nb[0] ? 0 : (blk < nb[0] + nb[1] ? 1 : 2);
    const uint32_t lb = blk - (c > 0 ? nb[0] : 0) - (c > 1 ? nb[1] : 0);
    const uint32_t bpr = plan.plane_w[c] / 8, by = lb / bpr, bx = lb - by * bpr;
    const unsigned gmask = 0xffu << ((tid & 31) & ~7);
    int* w8 = ws + (tid >> 3) * EB_STRIDE;
    // pass 1: thread t = row t of the block
    const uint8_t* src = planes + plan.plane_off[c] + (size_t)(by * 8 + t) * plan.plane_w[c] + bx * 8;
    const uint2 raw = *reinterpret_cast<const uint2*>(src);
    int d[8], o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        d[k] = (int)((raw.x >> (8 * k)) & 0xff) - 128;
        d[4 + k] = (int)((raw.y >> (8 * k)) & 0xff) - 128;
    }
    jfdct_1d<false>(d, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) w8[t * 8 + k] = o[k];  // row-major
    __syncwarp(gmask);
    // pass 2: thread t = column t
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = w8[k * 8 + t];
    jfdct_1d<true>(d, o);
    // quantise (jcdctmgr.c: the DCT output is scaled by 8, so the divisor is 8 * table entry; round half away from zero)
    int16_t* dst = coefs + (size_t)blk * 64;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int q = (int)plan.quant[c][k * 8 + t] << 3;
        int v = o[k];
        const bool neg = v < 0;
        v = ((neg ? -v : v) + (q >> 1)) / q;
        dst[k * 8 + t] = (int16_t)(neg ? -v : v);
    }
}

__global__ void __launch_bounds__(EB_PER_CTA * 8)
jpeg_fdct_kernel(const uint8_t* __restrict__ planes, JpegPlan plan, int16_t* __restrict__ coefs) {
    __shared__ int ws[EB_PER_CTA * EB_STRIDE];
    jpeg_fdct_body(planes, plan, coefs, ws);
}

__global__ void __launch_bounds__(EB_PER_CTA * 8)
jpeg_fdct_batch_kernel(const uint8_t* __restrict__ planes_base, const JpegEncJob* __restrict__ jobs, int16_t* __restrict__ coefs_base) {
    __shared__ int ws[EB_PER_CTA * EB_STRIDE];
    const JpegEncJob& J = jobs[blockIdx.y];
    jpeg_fdct_body(planes_base + J.planes_off, J.plan, coefs_base + J.coef_off, ws);
}

// `frames` jobs in two launches; max_* = the largest frame's extents
void launch_jpeg_encode_batch(const uint8_t* d_rgb_base, const JpegEncJob* d_jobs, int frames, uint32_t max_cw, uint32_t max_ch, uint32_t max_blocks,
                              uint8_t* d_planes_base, int16_t* d_coefs_base, cudaStream_t s) {
    if (frames <= 0) return;
    jpeg_enc_color_batch_kernel<<<dim3((max_cw + 31) / 32, (max_ch + 7) / 8, frames), 256, 0, s>>>(d_rgb_base, d_jobs, d_planes_base);
    jpeg_fdct_batch_kernel<<<dim3((max_blocks + EB_PER_CTA - 1) / EB_PER_CTA, frames), EB_PER_CTA * 8, 0, s>>>(d_planes_base, d_jobs, d_coefs_base);
}

void launch_jpeg_encode(const uint8_t* d_rgb, const JpegPlan& plan, uint8_t* d_planes, int16_t* d_coefs, cudaStream_t s) {
    dim3 g1((plan.plane_w[1] + 31) / 32, (plan.plane_h[1] + 7) / 8);
    jpeg_enc_color_kernel<<<g1, 256, 0, s>>>(d_rgb, plan, d_planes);
    const uint32_t nblk = (plan.plane_w[0] * plan.plane_h[0] + 2 * plan.plane_w[1] * plan.plane_h[1]) / 64;
    jpeg_fdct_kernel<<<(nblk + EB_PER_CTA - 1) / EB_PER_CTA, EB_PER_CTA * 8, 0, s>>>(d_planes, plan, d_coefs);
}

}  // namespace uf
