// kernels_prestem.cu — K1 + K2 + K3 in one kernel for frames at exactly twice the network size (640x480 -> RFB-320,
// 1280x960 -> RFB-640): `image::imageops::resize(.., Triangle)` (/root/reference/infer_server/src/nn.rs:74-80), the
// normalisation + HWC->NCHW of nn.rs:82-91 and the first Conv of the graph (3x3 stride 2, 3 -> 16, + BN + ReLU;
// nn.rs:181). The resized u8 frame and the normalised f32 tensor never exist in memory: a CTA reads the u8 source
// tile, resamples it in packed integer arithmetic, normalises through the 3x256 LUT into shared memory and convolves.
//
// Bit-exactness of the resize (the u8 values the convolution sees are the ones the reference would have produced):
// at exactly 2:1 every interior output has the taps 2o-1 .. 2o+2 with weights {1,3,3,1}/8, exact binary fractions, so
// every product and partial sum of the reference's f32 arithmetic is exact and the result is the integer
// (sum_ij w_i w_j p_ij + 32) >> 6 (see resize_half_exact_kernel in kernels_preproc.cu, which this kernel restates on
// 2-D tiles). First / last output row and column have clamped, renormalised taps (not binary fractions): they are
// recomputed in f32 from the tap tables in the reference's operation order. `dbg_resized` (parity hook) receives the
// u8 values the kernel convolved.
#include "kernels.h"
#include "pdl.cuh"

namespace uf {

constexpr int PS_TX = 32, PS_TY = 8;                  // stem outputs per CTA (thread = one output pixel, 16 channels)
constexpr int PS_RW = 2 * PS_TX + 1, PS_RH = 2 * PS_TY + 1;  // resized pixels under the tile (3x3 stride 2, pad 1)
constexpr int PS_SCOLS = 4 * PS_TX + 8;               // source columns staged per row, aligned down to a multiple of 4
constexpr int PS_SB = PS_SCOLS * 3;                   // bytes per staged source row (multiple of 4)
constexpr int PS_WORDS = PS_SB / 4;
constexpr int PS_HALF = PS_TX + 1;                    // resized columns per parity

struct PrestemWeights {
    float w[27 * 16];  // [ky][kx][ci][co]
    float b[16];
};

__device__ __forceinline__ uint8_t prestem_px_f32(const uint8_t* __restrict__ sf, int sw, const ResizeTapsDev& t, int oy, int ox,
                                                  int ch, int round_intermediate) {
    const int vl = t.vleft[oy], vn = t.vn[oy], hl = t.hleft[ox], hn = t.hn[ox];
    const float* vw = t.vw + (size_t)oy * t.vmax;
    const float* hw = t.hw + (size_t)ox * t.hmax;
    float acc = 0.f;
    for (int j = 0; j < hn; ++j) {
        float v = 0.f;
        for (int i = 0; i < vn; ++i)
            v = __fadd_rn(v, __fmul_rn((float)sf[((size_t)(vl + i) * sw + hl + j) * 3 + ch], vw[i]));
        if (round_intermediate) v = roundf(fminf(fmaxf(v, 0.f), 255.f));
        acc = __fadd_rn(acc, __fmul_rn(v, hw[j]));
    }
    acc = acc < 0.f ? 0.f : (acc > 255.f ? 255.f : acc);
    return (uint8_t)roundf(acc);
}

constexpr int PS_THREADS = 128;   // thread = output column x 2 output rows (ty, ty + PS_TY / 2): every weight fetched feeds 2 FFMAs
constexpr int PS_WARPS = PS_THREADS / 32;

__global__ void __launch_bounds__(PS_THREADS)
resize2_stem_kernel(U8View src, const float* __restrict__ lut, ResizeTapsDev t, TView out, const __grid_constant__ PrestemWeights wts,
                    int relu, int round_intermediate, uint8_t* __restrict__ dbg_resized) {
    __shared__ __align__(16) uint16_t vt[PS_RH][PS_SB];          // vertical sums a + 3b + 3c + d per source byte column
    __shared__ float s_in[3][PS_RH][2][PS_HALF];                 // normalised resized tile: [channel][row][column parity][column / 2]
    __shared__ float s_lut[768];
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 768; i += PS_THREADS) s_lut[i] = lut[i];  // static table: may be read before the wait
    pdl_wait();
    const int sw = src.W, sh = src.H, dw = sw >> 1, dh = sh >> 1;
    const int ox0 = blockIdx.x * PS_TX, oy0 = blockIdx.y * PS_TY, n = blockIdx.z;
    const int rx0 = 2 * ox0 - 1, ry0 = 2 * oy0 - 1;  // first resized column / row under the tile (may be -1: padding)
    const int sx0 = 4 * ox0 - 4;                     // first staged source column (16-byte aligned in the row)
    const uint8_t* sf = src.p + (size_t)n * src.frame_stride;
    const int row_words = (sw * 3) >> 2;
    const int word0 = (sx0 * 3) >> 2;                // may be negative at the left edge: clamped below
    const int row_bytes = sw * 3;

    // 1. vertical pass, packed 16-bit lanes: thread = one 4-byte column of the source tile, walking down its 2 * PS_RH + 2
    //    source rows with a sliding window — every source word is loaded and unpacked once and feeds two resized rows
    //    (rows 2r-1 .. 2r+2 for resized row r). Row indices are clamped into the frame one by one: only padding rows
    //    (zeroed in pass 2) and border rows (recomputed in pass 3) see the difference.
    if (tid < PS_WORDS) {
        const int wd = min(max(word0 + tid, 0), row_words - 1);
        const unsigned* col = reinterpret_cast<const unsigned*>(sf) + wd;
        const int ys0 = 2 * ry0 - 1;  // source row of the first tap of the first resized row
        auto ld = [&](int k) { return __ldg(col + (size_t)min(max(ys0 + k, 0), sh - 1) * row_words); };
        unsigned v0 = ld(0), v1 = ld(1);
        unsigned e0 = v0 & 0x00ff00ffu, o0 = (v0 >> 8) & 0x00ff00ffu, e1 = v1 & 0x00ff00ffu, o1 = (v1 >> 8) & 0x00ff00ffu;
#pragma unroll
        for (int rl = 0; rl < PS_RH; ++rl) {
            const unsigned v2 = ld(2 * rl + 2), v3 = ld(2 * rl + 3);
            const unsigned e2 = v2 & 0x00ff00ffu, o2 = (v2 >> 8) & 0x00ff00ffu, e3 = v3 & 0x00ff00ffu, o3 = (v3 >> 8) & 0x00ff00ffu;
            // bytes 0,2 and bytes 1,3 as packed 16-bit lanes; sums stay below 2^11 (2^8 after the optional rounding)
            unsigned e = e0 + e3 + 3u * (e1 + e2);
            unsigned o = o0 + o3 + 3u * (o1 + o2);
            if (round_intermediate) {
                e = ((e + 0x00040004u) >> 3) & 0x00ff00ffu;
                o = ((o + 0x00040004u) >> 3) & 0x00ff00ffu;
            }
            *reinterpret_cast<uint2*>(&vt[rl][4 * tid]) = make_uint2(__byte_perm(e, o, 0x5410), __byte_perm(e, o, 0x7632));  // u16 order: byte 0,1,2,3
            e0 = e2; o0 = o2; e1 = e3; o1 = o3;
        }
    }
    __syncthreads();

    // 2. horizontal pass + normalise: one resized pixel (3 channels) per item; the 12 sums of a pixel sit in 7 consecutive
    //    32-bit words. Padding (outside the resized frame) = 0 in normalised space, as in the ONNX Conv.
    const int sh6 = round_intermediate ? 3 : 6, half = round_intermediate ? 4 : 32;
    uint8_t* dbg = dbg_resized ? dbg_resized + (size_t)n * dw * dh * 3 : nullptr;
    for (int it = tid; it < PS_RH * PS_RW; it += PS_THREADS) {
        const int rl = it / PS_RW, cl = it - rl * PS_RW;  // division by a constant
        const int rx = rx0 + cl, ry = ry0 + rl;
        const bool inside = ry >= 0 && ry < dh && rx >= 0 && rx < dw;
        // u16 index 6 cl + 3 .. 6 cl + 14 (source columns 2 rx - 1 .. 2 rx + 2, three channels each)
        const unsigned* x = reinterpret_cast<const unsigned*>(&vt[rl][0]) + 3 * cl + 1;
        const unsigned w0 = x[0], w1 = x[1], w2 = x[2], w3 = x[3], w4 = x[4], w5 = x[5], w6 = x[6];
        const int r4 = (int)(w0 >> 16) + (int)(w5 & 0xffffu) + 3 * ((int)(w2 & 0xffffu) + (int)(w3 >> 16));
        const int g4 = (int)(w1 & 0xffffu) + (int)(w5 >> 16) + 3 * ((int)(w2 >> 16) + (int)(w4 & 0xffffu));
        const int b4 = (int)(w1 >> 16) + (int)(w6 & 0xffffu) + 3 * ((int)(w3 & 0xffffu) + (int)(w4 >> 16));
        const int ur = (r4 + half) >> sh6, ug = (g4 + half) >> sh6, ub = (b4 + half) >> sh6;
        float* dst = &s_in[0][rl][cl & 1][cl >> 1];
        dst[0] = inside ? s_lut[ur] : 0.0f;
        dst[PS_RH * 2 * PS_HALF] = inside ? s_lut[256 + ug] : 0.0f;
        dst[2 * PS_RH * 2 * PS_HALF] = inside ? s_lut[512 + ub] : 0.0f;
        if (dbg && inside && cl >= 1 && rl >= 1) {
            uint8_t* q = dbg + ((size_t)ry * dw + rx) * 3;
            q[0] = (uint8_t)ur; q[1] = (uint8_t)ug; q[2] = (uint8_t)ub;
        }
    }
    // 3. border outputs of the resize in f32 from the tap tables (clamped, renormalised taps are not binary fractions):
    //    resized rows 0 / dh - 1 and columns 0 / dw - 1 where the tile holds them
    const bool edge = ry0 <= 0 || ry0 + PS_RH >= dh || rx0 <= 0 || rx0 + PS_RW >= dw;  // CTA-uniform
    if (edge) {
        __syncthreads();
        const int br0 = 0 - ry0, br1 = dh - 1 - ry0;  // local rows of the two border rows (maybe outside the tile)
        const int bc0 = 0 - rx0, bc1 = dw - 1 - rx0;
        // items: [2 border rows x PS_RW columns] then [2 border columns x PS_RH rows], 3 channels each
        for (int it = tid; it < (2 * PS_RW + 2 * PS_RH) * 3; it += PS_THREADS) {
            const int ch = it % 3, q = it / 3;
            int rl, cl;
            if (q < 2 * PS_RW) { rl = q < PS_RW ? br0 : br1; cl = q < PS_RW ? q : q - PS_RW; }
            else { const int q2 = q - 2 * PS_RW; cl = q2 < PS_RH ? bc0 : bc1; rl = q2 < PS_RH ? q2 : q2 - PS_RH; }
            if (rl < 0 || rl >= PS_RH || cl < 0 || cl >= PS_RW) continue;
            const int rx = rx0 + cl, ry = ry0 + rl;
            if (rx < 0 || rx >= dw || ry < 0 || ry >= dh) continue;
            const uint8_t u = prestem_px_f32(sf, sw, t, ry, rx, ch, round_intermediate);
            s_in[ch][rl][cl & 1][cl >> 1] = s_lut[ch * 256 + u];
            if (dbg && cl >= 1 && rl >= 1) dbg[((size_t)ry * dw + rx) * 3 + ch] = u;
        }
    }
    __syncthreads();

    // 4. the convolution: thread = output column x rows (ty, ty + 4), weights as constant-bank operands shared by both
    const int tx = lane, ty = warp;
    const int ox = ox0 + tx;
    if (ox >= out.W) return;
    float acc0[16], acc1[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc0[c] = acc1[c] = wts.b[c];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float v0 = s_in[ci][2 * ty + ky][kx & 1][tx + (kx >> 1)];
                const float v1 = s_in[ci][2 * (ty + PS_TY / 2) + ky][kx & 1][tx + (kx >> 1)];
#pragma unroll
                for (int co = 0; co < 16; ++co) {
                    const float wv = wts.w[((ky * 3 + kx) * 3 + ci) * 16 + co];
                    acc0[co] = fmaf(v0, wv, acc0[co]);
                    acc1[co] = fmaf(v1, wv, acc1[co]);
                }
            }
#pragma unroll
    for (int half_i = 0; half_i < 2; ++half_i) {
        const int oy = oy0 + ty + half_i * (PS_TY / 2);
        if (oy >= out.H) continue;
        const float* acc = half_i ? acc1 : acc0;
        float* op = out.p + (size_t)n * out.frame_stride + ((size_t)oy * out.W + ox) * out.pix_stride;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4*>(op + q * 4) = v;
        }
    }
}

bool prestem_supported(const uint8_t* src, long long src_frame_stride, int sw, int sh, int net_w, int net_h, const ResizeTapsDev& t) {
    return sw == 2 * net_w && sh == 2 * net_h && sw % 4 == 0 && net_w >= 4 && net_h >= 2 && (reinterpret_cast<size_t>(src) & 3) == 0 &&
           src_frame_stride % 4 == 0 && t.vmax == 4 && t.hmax == 4;
}

// host_w: 27*16 weights [ky][kx][ci][co] followed by 16 biases, in HOST memory; src = frames at twice the network size
void launch_prestem(const U8View& src, const float* lut, const ResizeTapsDev& t, const TView& out, const float* host_w, int relu,
                    int round_intermediate, int frames, uint8_t* dbg_resized, cudaStream_t s) {
    dim3 grid((out.W + PS_TX - 1) / PS_TX, (out.H + PS_TY - 1) / PS_TY, frames);
    launch_pdl(resize2_stem_kernel, grid, dim3(PS_THREADS), 0, s, src, lut, t, out, *reinterpret_cast<const PrestemWeights*>(host_w),
               relu, round_intermediate, dbg_resized);
}

}  // namespace uf
