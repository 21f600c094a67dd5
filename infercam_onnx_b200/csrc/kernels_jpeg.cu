// kernels_jpeg.cu — N2: the sample-domain half of JPEG decoding on the GPU (see jpeg_decode.h). Replaces the CPU work of
// `turbojpeg::decompress_image` (/root/reference/infer_server/src/inferer.rs:35) after Huffman decoding, and restates
// libjpeg-turbo's DEFAULT algorithms in exact integer arithmetic, so the RGB bytes are the ones the reference's decoder
// produces:
//   jpeg_idct_kernel   dequantisation + `jpeg_idct_islow` (jidctint.c: 13-bit constants, two passes, DESCALE rounding,
//                      range-limit table incl. its wrap-around) -> one u8 plane per component
//   jpeg_color_kernel  `h2v1_fancy_upsample` / `h2v2_fancy_upsample` (jdsample.c; edge rows replicate the first / last REAL
//                      row, jdmainct.c) + `ycc_rgb_convert` (jdcolor.c fixed-point tables, computed on the fly) -> RGB8 HWC
// Both are HBM/issue-light (a few hundred integer ops per 8x8 block) next to the CNN that follows.
#include "jpeg_decode.h"
#include "kernels.h"

namespace uf {

constexpr int JB_PER_CTA = 32;     // 8x8 blocks per CTA, 8 threads each
constexpr int JB_STRIDE = 72;      // words per block in shared memory (64 + 8: the 4 blocks of a warp hit distinct banks)

#define JFIX_0_298631336 2446
#define JFIX_0_390180644 3196
#define JFIX_0_541196100 4433
#define JFIX_0_765366865 6270
#define JFIX_0_899976223 7373
#define JFIX_1_175875602 9633
#define JFIX_1_501321110 12299
#define JFIX_1_847759065 15137
#define JFIX_1_961570560 16069
#define JFIX_2_053119869 16819
#define JFIX_2_562915447 20995
#define JFIX_3_072711026 25172

__device__ __forceinline__ int jdescale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// range_limit[x & RANGE_MASK] of jidctint.c (sample_range_limit + CENTERJSAMPLE): clamp(x + 128, 0, 255) for sane
// values, libjpeg's wrap-around for wild ones
__device__ __forceinline__ unsigned jidct_limit(int x) {
    const int v = x & 1023;
    return v < 128 ? v + 128 : v < 512 ? 255 : v < 896 ? 0 : v - 896;
}

// one 1-D pass of jpeg_idct_islow on 8 values (already dequantised / from the workspace); `shift` = final DESCALE
__device__ __forceinline__ void jidct_1d(const int in[8], int out[8], int shift) {
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * JFIX_0_541196100;
    int tmp2 = z1 + z3 * (-JFIX_1_847759065);
    int tmp3 = z1 + z2 * JFIX_0_765366865;
    int tmp0 = (in[0] + in[4]) << 13;
    int tmp1 = (in[0] - in[4]) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * JFIX_1_175875602;
    tmp0 *= JFIX_0_298631336; tmp1 *= JFIX_2_053119869; tmp2 *= JFIX_3_072711026; tmp3 *= JFIX_1_501321110;
    z1 *= -JFIX_0_899976223; z2 *= -JFIX_2_562915447; z3 *= -JFIX_1_961570560; z4 *= -JFIX_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    out[0] = jdescale(tmp10 + tmp3, shift); out[7] = jdescale(tmp10 - tmp3, shift);
    out[1] = jdescale(tmp11 + tmp2, shift); out[6] = jdescale(tmp11 - tmp2, shift);
    out[2] = jdescale(tmp12 + tmp1, shift); out[5] = jdescale(tmp12 - tmp1, shift);
    out[3] = jdescale(tmp13 + tmp0, shift); out[4] = jdescale(tmp13 - tmp0, shift);
}

__global__ void __launch_bounds__(JB_PER_CTA * 8)
jpeg_idct_kernel(JpegBatchDev b) {
    __shared__ int coef[JB_PER_CTA * JB_STRIDE];
    __shared__ JpegPlan plan;
    const int tid = threadIdx.x;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(b.plans + blockIdx.y);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&plan);
        for (int i = tid; i < (int)(sizeof(JpegPlan) / 4); i += JB_PER_CTA * 8) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t blk = blockIdx.x * JB_PER_CTA + (tid >> 3);
    if (blk >= plan.nblocks) return;  // whole 8-thread groups leave together; no block-wide barrier follows
    if (b.status && b.status[blockIdx.y] != 0) return;  // the device Huffman decoder declined the frame: its lists are not valid
    const int t = tid & 7;
    const unsigned gmask = 0xffu << ((tid & 31) & ~7);  // the 8 lanes working on this block
    int* cf = coef + (tid >> 3) * JB_STRIDE;
    const uint32_t mcu = blk / plan.blocks_per_mcu, sl = blk - mcu * plan.blocks_per_mcu;
    const int c = plan.slot_comp[sl];
    const uint32_t my = mcu / plan.mcus_x, mx = mcu - my * plan.mcus_x;
    const uint32_t bx = mx * plan.hs[c] + plan.slot_h[sl], by = my * plan.vs[c] + plan.slot_v[sl];
    // 1. the block's coefficients, dequantised (DEQUANTIZE: coefficient * quantval)
    // zero, then scatter the list of nonzeros (host Huffman path: DC included; device path: AC only, DC from its own array)
#pragma unroll
    for (int i = 0; i < 8; ++i) cf[i * 8 + t] = 0;
    __syncwarp(gmask);
    {
        const uint32_t* offs = b.offs + plan.offs_base + blk;
        const uint32_t e0 = offs[0], e1 = min(offs[1], e0 + 64u);
        const uint32_t* ent = b.entries + plan.entries_base;
        for (uint32_t e = e0 + t; e < e1; e += 8) {
            const uint32_t v = ent[e];
            const int idx = (v >> 16) & 63;
            cf[idx] = (int)(short)(v & 0xffffu) * (int)plan.quant[c][idx];
        }
        if (b.dcv && t == 0) cf[0] = (int)b.dcv[plan.offs_base + blk] * (int)plan.quant[c][0];
    }
    __syncwarp(gmask);
    // 2. columns: thread t = column t; results stored transposed (index column * 8 + row) so that pass 2 reads conflict-free
    int in[8], ws[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) in[r] = cf[r * 8 + t];
    jidct_1d(in, ws, 13 - 2);
    __syncwarp(gmask);
#pragma unroll
    for (int r = 0; r < 8; ++r) cf[t * 8 + r] = ws[r];
    __syncwarp(gmask);
    // 3. rows: thread t = row t
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) in[cc] = cf[cc * 8 + t];
    jidct_1d(in, ws, 13 + 2 + 3);
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lo |= jidct_limit(ws[k]) << (8 * k);
        hi |= jidct_limit(ws[4 + k]) << (8 * k);
    }
    const size_t frame_planes = ((size_t)plan.planes_off_hi << 32) | plan.planes_off_lo;
    uint8_t* dst = b.planes + frame_planes + plan.plane_off[c] + (size_t)(by * 8 + t) * plan.plane_w[c] + bx * 8;
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
}

__device__ __forceinline__ unsigned jclamp(int v) { return (unsigned)min(max(v, 0), 255); }

// jdcolor.c: Cr_r_tab[cr] = (FIX(1.40200) * (cr-128) + ONE_HALF) >> 16, Cb_b_tab likewise with 1.77200,
// G = y + ((-FIX(0.34414) * (cb-128) + ONE_HALF - FIX(0.71414) * (cr-128)) >> 16); arithmetic shifts
__device__ __forceinline__ void jycc(int y, int cb, int cr, unsigned& r, unsigned& g, unsigned& bl) {
    cb -= 128; cr -= 128;
    r = jclamp(y + ((91881 * cr + 32768) >> 16));
    g = jclamp(y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
    bl = jclamp(y + ((116130 * cb + 32768) >> 16));
}

// The chroma samples of a quad of pixels x0 .. x0 + 3 (x0 a multiple of 4) of row y at once: the four chroma columns it
// touches are fetched once (jup would fetch them per pixel). Same arithmetic as jup, pixel for pixel.
__device__ __forceinline__ void jup_quad(const uint8_t* __restrict__ plane, int pw, int rw, int rh, int hr, int vr, int x0, int y, int nx, int out[4]) {
    if (hr == 1) {
        const uint8_t* in = plane + (size_t)y * pw + x0;
#pragma unroll
        for (int k = 0; k < 4; ++k) out[k] = k < nx ? in[k] : 0;
        return;
    }
    const int c0 = x0 >> 1;  // pixels 0, 1 -> chroma column c0; pixels 2, 3 -> c0 + 1
    const bool has_a = c0 > 0, has_c = c0 + 1 < rw, has_d = c0 + 2 < rw;
    if (vr == 1) {  // h2v1_fancy_upsample
        const uint8_t* in = plane + (size_t)y * pw + c0;
        const int a = has_a ? in[-1] : 0, bq = in[0], c = has_c ? in[1] : 0, d = has_d ? in[2] : 0;
        out[0] = has_a ? (bq * 3 + a + 1) >> 2 : bq;
        out[1] = has_c ? (bq * 3 + c + 2) >> 2 : bq;       // (c0 == rw - 1: the last column repeats)
        out[2] = (c * 3 + bq + 1) >> 2;
        out[3] = has_d ? (c * 3 + d + 2) >> 2 : c;
        return;
    }
    // h2v2_fancy_upsample: the nearer input row counts 3, the farther 1; context rows replicate the first / last real row
    const int cy = y >> 1;
    const int oy = min(max((y & 1) ? cy + 1 : cy - 1, 0), rh - 1);
    const uint8_t* in0 = plane + (size_t)cy * pw + c0;
    const uint8_t* in1 = plane + (size_t)oy * pw + c0;
    const int ta = has_a ? in0[-1] * 3 + in1[-1] : 0, tb = in0[0] * 3 + in1[0], tc = has_c ? in0[1] * 3 + in1[1] : 0,
              td = has_d ? in0[2] * 3 + in1[2] : 0;
    out[0] = has_a ? (tb * 3 + ta + 8) >> 4 : (tb * 4 + 8) >> 4;
    out[1] = has_c ? (tb * 3 + tc + 7) >> 4 : (tb * 4 + 7) >> 4;
    out[2] = (tc * 3 + tb + 8) >> 4;
    out[3] = has_d ? (tc * 3 + td + 7) >> 4 : (tc * 4 + 7) >> 4;
}

__global__ void __launch_bounds__(256)
jpeg_color_kernel(JpegBatchDev b) {
    const JpegPlan& plan = b.plans[blockIdx.y];  // (a few fields, the same for every thread: served from L1)
    if (b.status && b.status[blockIdx.y] != 0) return;  // declined by the device Huffman decoder: redone on the host
    const int w = (int)plan.w, h = (int)plan.h;
    const int quads = (w + 3) >> 2;
    const uint32_t q = blockIdx.x * 256u + threadIdx.x;  // (frames are at most 16384 x 16384: quads * h < 2^32)
    if (q >= (uint32_t)quads * (uint32_t)h) return;
    const int y = (int)(q / (uint32_t)quads), x0 = (int)(q - (uint32_t)y * (uint32_t)quads) * 4;
    const uint8_t* planes = b.planes + (((size_t)plan.planes_off_hi << 32) | plan.planes_off_lo);
    uint8_t* rgb = b.rgb + (((size_t)plan.rgb_off_hi << 32) | plan.rgb_off_lo) + ((size_t)y * w + x0) * 3;
    const uint8_t* yp = planes + plan.plane_off[0] + (size_t)y * plan.plane_w[0] + x0;
    unsigned px[12];
    const int nx = min(4, w - x0);
    const uchar4 y4 = *reinterpret_cast<const uchar4*>(yp);  // (x0 and the plane's width and offset are multiples of 4 or more)
    const int ys[4] = {y4.x, y4.y, y4.z, y4.w};
    if (plan.ncomp == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) px[3 * k] = px[3 * k + 1] = px[3 * k + 2] = k < nx ? ys[k] : 0;
    } else {
        const int hr = (int)(plan.hmax / plan.hs[1]), vr = (int)(plan.vmax / plan.vs[1]);
        const int pw = (int)plan.plane_w[1], rw = (int)plan.real_w[1], rh = (int)plan.real_h[1];
        int cb[4], cr[4];
        jup_quad(planes + plan.plane_off[1], pw, rw, rh, hr, vr, x0, y, nx, cb);
        jup_quad(planes + plan.plane_off[2], pw, rw, rh, hr, vr, x0, y, nx, cr);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < nx) jycc(ys[k], cb[k], cr[k], px[3 * k], px[3 * k + 1], px[3 * k + 2]);
            else px[3 * k] = px[3 * k + 1] = px[3 * k + 2] = 0;
        }
    }
    if (nx == 4 && (reinterpret_cast<size_t>(rgb) & 3) == 0) {
        unsigned* o = reinterpret_cast<unsigned*>(rgb);
        o[0] = px[0] | px[1] << 8 | px[2] << 16 | px[3] << 24;
        o[1] = px[4] | px[5] << 8 | px[6] << 16 | px[7] << 24;
        o[2] = px[8] | px[9] << 8 | px[10] << 16 | px[11] << 24;
    } else {
        for (int k = 0; k < nx * 3; ++k) rgb[k] = (uint8_t)px[k];
    }
}

void launch_jpeg_decode(const JpegBatchDev& b, int frames, uint32_t max_nblocks, uint32_t max_w, uint32_t max_h, cudaStream_t s) {
    if (frames <= 0) return;
    jpeg_idct_kernel<<<dim3((max_nblocks + JB_PER_CTA - 1) / JB_PER_CTA, frames), JB_PER_CTA * 8, 0, s>>>(b);
    const long long quads = (long long)((max_w + 3) / 4) * max_h;
    jpeg_color_kernel<<<dim3((unsigned)((quads + 255) / 256), frames), 256, 0, s>>>(b);
}

}  // namespace uf
