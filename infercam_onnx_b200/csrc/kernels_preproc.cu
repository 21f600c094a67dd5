// kernels_preproc.cu — K1 (bit-exact triangle resize) and K2 (normalise hook).
//
// K1 replaces `image::imageops::resize(input, W, H, FilterType::Triangle)`
// (/root/reference/infer_server/src/nn.rs:74-80; algorithm: image 0.24.5
// src/imageops/sample.rs vertical_sample + horizontal_sample). The u8 output must be
// bit-exact, so every product and sum is a separate IEEE round-to-nearest operation
// (__fmul_rn/__fadd_rn are never contracted into FMA), taps are accumulated in ascending
// order, the vertical pass result stays f32 in shared memory (image 0.24.x keeps an
// Rgba32FImage between the passes) and the final value is clamp -> round-half-away -> u8.
//
// One CTA produces a tile_h x tile_w tile of the destination for one frame:
//   pass 1 (vertical):   tmp[r][cb] = sum_i src[vleft[oy]+i][col0*3+cb] * vw[oy][i]   -> smem f32
//   pass 2 (horizontal): dst[oy][ox][c] = round(clamp(sum_i tmp[r][(l+i)*3+c] * hw[ox][i]))
// Threads walk bytes of a row (HWC => contiguous) so global loads and stores coalesce.
#include "kernels.h"

namespace uf {

__global__ void __launch_bounds__(256)
resize_triangle_kernel(const uint8_t* __restrict__ src, long long src_frame_stride, int sw, int sh,
                       uint8_t* __restrict__ dst, long long dst_frame_stride, int dw, int dh,
                       ResizeTapsDev t, int round_intermediate) {
    extern __shared__ float tmp_s[];  // tile_h x (max_cols*3)
    const int ox0 = blockIdx.x * t.tile_w, oy0 = blockIdx.y * t.tile_h;
    const int ox1 = min(ox0 + t.tile_w, dw), oy1 = min(oy0 + t.tile_h, dh);
    const int rows = oy1 - oy0, tw = ox1 - ox0;
    const int col0 = t.hleft[ox0];
    const int col1 = t.hleft[ox1 - 1] + t.hn[ox1 - 1];
    const int nc3 = (col1 - col0) * 3;
    const int pitch = t.max_cols * 3;
    const uint8_t* s = src + (size_t)blockIdx.z * src_frame_stride + (size_t)col0 * 3;
    const size_t row_bytes = (size_t)sw * 3;

    for (int idx = threadIdx.x; idx < rows * nc3; idx += blockDim.x) {
        const int r = idx / nc3, cb = idx - r * nc3;
        const int oy = oy0 + r;
        const int l = t.vleft[oy], n = t.vn[oy];
        const float* w = t.vw + (size_t)oy * t.vmax;
        const uint8_t* sp = s + (size_t)l * row_bytes + cb;
        float acc = 0.0f;
        for (int i = 0; i < n; ++i) {
            acc = __fadd_rn(acc, __fmul_rn((float)sp[(size_t)i * row_bytes], w[i]));
        }
        if (round_intermediate) acc = roundf(fminf(fmaxf(acc, 0.0f), 255.0f));
        tmp_s[r * pitch + cb] = acc;
    }
    __syncthreads();
    uint8_t* d = dst + (size_t)blockIdx.z * dst_frame_stride;
    const int tw3 = tw * 3;
    for (int idx = threadIdx.x; idx < rows * tw3; idx += blockDim.x) {
        const int r = idx / tw3, rem = idx - r * tw3;
        const int oxl = rem / 3, c = rem - oxl * 3;
        const int ox = ox0 + oxl;
        const int l = t.hleft[ox] - col0, n = t.hn[ox];
        const float* w = t.hw + (size_t)ox * t.hmax;
        const float* tp = tmp_s + r * pitch + l * 3 + c;
        float acc = 0.0f;
        for (int i = 0; i < n; ++i) {
            acc = __fadd_rn(acc, __fmul_rn(tp[i * 3], w[i]));
        }
        // sample.rs: clamp(t, 0, 255) then FloatNearest (f32::round, half away from zero)
        acc = acc < 0.0f ? 0.0f : (acc > 255.0f ? 255.0f : acc);
        d[((size_t)(oy0 + r) * dw + ox) * 3 + c] = (uint8_t)roundf(acc);
    }
}

void launch_resize(const uint8_t* src, long long src_frame_stride, int sw, int sh, uint8_t* dst,
                   long long dst_frame_stride, int dw, int dh, int frames, const ResizeTapsDev& t,
                   int round_intermediate, cudaStream_t s) {
    dim3 grid((dw + t.tile_w - 1) / t.tile_w, (dh + t.tile_h - 1) / t.tile_h, frames);
    size_t smem = (size_t)t.tile_h * t.max_cols * 3 * sizeof(float);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(resize_triangle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    resize_triangle_kernel<<<grid, 256, smem, s>>>(src, src_frame_stride, sw, sh, dst, dst_frame_stride, dw,
                                                   dh, t, round_intermediate);
}

// K2 hook: out[c][y][x] = lut[c][hwc[y][x][c]]; the LUT holds (v/255 - mean[c]) / std[c]
// evaluated on the host with the reference's f32 operation order (nn.rs:85-88), so the
// tensor is bit-exact by construction (only 3 x 256 distinct values exist).
__global__ void normalise_nchw_kernel(const uint8_t* __restrict__ hwc, int w, int h,
                                      const float* __restrict__ lut, float* __restrict__ out) {
    const int total = w * h * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i / (w * h), pix = i - c * (w * h);
        out[i] = lut[c * 256 + hwc[(size_t)pix * 3 + c]];
    }
}

void launch_normalise_nchw(const uint8_t* hwc, int w, int h, const float* lut, float* out, cudaStream_t s) {
    int total = w * h * 3;
    normalise_nchw_kernel<<<(total + 255) / 256, 256, 0, s>>>(hwc, w, h, lut, out);
}

}  // namespace uf
