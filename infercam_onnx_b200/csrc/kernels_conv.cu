// kernels_conv.cu — K3..K7: the UltraFace conv stack (replaces tract's SimplePlan::run,
// /root/reference/infer_server/src/nn.rs:181). NHWC fp32 activations, fp32 accumulate
// (BASELINE.json north_star: raw tensors within 1e-4 abs of the reference => no tf32/bf16 here).
// Bias, folded BatchNorm, residual add and ReLU live in the epilogue of the producing conv.
#include "kernels.h"
#include "pdl.cuh"

namespace uf {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------------
// Generic direct convolution: one thread per output element, co fastest (coalesced stores,
// coalesced weight reads from [ky][kx][ci][co]). Correct for every Conv the lowering accepts;
// used for the shapes without a specialised kernel (e.g. the 256->6/12 3x3 heads on the 4x5 map)
// and as the in-library cross-check (UF_FLAG_FORCE_GENERIC).
// ---------------------------------------------------------------------------------------------
template <bool U8IN>
__global__ void __launch_bounds__(256)
conv_generic_kernel(TView in, U8View in8, const float* __restrict__ lut, TView out, TView res, int has_res,
                    const float* __restrict__ w, const float* __restrict__ b, ConvParams p, long long total) {
    const int cin_g = p.cin / p.groups, cout_g = p.cout / p.groups;
    const int Hi = U8IN ? in8.H : in.H, Wi = U8IN ? in8.W : in.W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(idx % p.cout);
        long long pix = idx / p.cout;
        const int x = (int)(pix % out.W);
        pix /= out.W;
        const int y = (int)(pix % out.H);
        const int n = (int)(pix / out.H);
        const int g = co / cout_g;
        float acc = b[co];
        for (int ky = 0; ky < p.k; ++ky) {
            const int iy = y * p.stride - p.pad + ky * p.dil;
            if (iy < 0 || iy >= Hi) continue;
            for (int kx = 0; kx < p.k; ++kx) {
                const int ix = x * p.stride - p.pad + kx * p.dil;
                if (ix < 0 || ix >= Wi) continue;
                const float* wp = w + ((size_t)(ky * p.k + kx) * cin_g) * p.cout + co;
                if (U8IN) {
                    const uint8_t* ip = in8.p + (size_t)n * in8.frame_stride + ((size_t)iy * Wi + ix) * 3 + g * cin_g;
                    for (int ci = 0; ci < cin_g; ++ci)
                        acc = fmaf(lut[(g * cin_g + ci) * 256 + ip[ci]], wp[(size_t)ci * p.cout], acc);
                } else {
                    const float* ip = in.p + (size_t)n * in.frame_stride + ((size_t)iy * Wi + ix) * in.pix_stride + g * cin_g;
                    for (int ci = 0; ci < cin_g; ++ci) acc = fmaf(ip[ci], wp[(size_t)ci * p.cout], acc);
                }
            }
        }
        const size_t opix = (size_t)y * out.W + x;
        if (has_res) acc += res.p[(size_t)n * res.frame_stride + opix * res.pix_stride + co];
        if (p.relu) acc = fmaxf(acc, 0.0f);
        out.p[(size_t)n * out.frame_stride + opix * out.pix_stride + co] = acc;
    }
}

static int grid_for(long long total, int block, int per_sm = 16) {
    long long g = (total + block - 1) / block;
    long long cap = 148LL * per_sm;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

void launch_conv_generic(const TView& in, const U8View* in_u8, const float* lut, const TView& out,
                         const TView* res, const float* w_kkio, const float* b, const ConvParams& p,
                         int frames, cudaStream_t s) {
    long long total = (long long)frames * out.H * out.W * p.cout;
    TView r = res ? *res : TView{};
    if (in_u8)
        conv_generic_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(in, *in_u8, lut, out, r, res != nullptr, w_kkio, b, p, total);
    else
        conv_generic_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(in, U8View{}, lut, out, r, res != nullptr, w_kkio, b, p, total);
}

// ---------------------------------------------------------------------------------------------
// K3 stem: 3x3 stride 2 pad 1, 3 -> 16 channels, input = resized u8 frame normalised through
// the LUT while it is staged (K2 fused: the f32 NCHW tensor of nn.rs:82-91 is never written).
// CTA = 16x16 output pixels; 33x33x3 normalised inputs + weights in shared memory; each thread
// keeps 16 accumulators and writes its pixel's 64 contiguous bytes.
// ---------------------------------------------------------------------------------------------
constexpr int STEM_T = 16;
constexpr int STEM_IN = 2 * STEM_T + 1;

// weights travel in the kernel parameter space: every FFMA takes its weight as a constant-bank operand
// (broadcasting them from shared memory costs one smem wavefront per FFMA and caps the FMA pipe at 25 %)
struct StemWeights {
    float w[27 * 16];  // [ky][kx][ci][co]
    float b[16];
};

__global__ void __launch_bounds__(STEM_T * STEM_T)
stem_kernel(U8View in, const float* __restrict__ lut, TView out, const __grid_constant__ StemWeights wts, int relu) {
    __shared__ float s_in[STEM_IN * STEM_IN * 3];
    __shared__ float s_lut[768];
    pdl_launch_dependents();
    const int tid = threadIdx.x;
    for (int i = tid; i < 768; i += blockDim.x) s_lut[i] = lut[i];  // static table: may be read before the wait
    pdl_wait();
    __syncthreads();
    const int ox0 = blockIdx.x * STEM_T, oy0 = blockIdx.y * STEM_T, n = blockIdx.z;
    const int ix0 = ox0 * 2 - 1, iy0 = oy0 * 2 - 1;
    const uint8_t* ip = in.p + (size_t)n * in.frame_stride;
    for (int i = tid; i < STEM_IN * STEM_IN; i += blockDim.x) {
        const int py = i / STEM_IN, px = i - py * STEM_IN;  // division by a constant
        const int ix = ix0 + px, iy = iy0 + py;
        float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;  // zero padding is applied in normalised space, as in the ONNX Conv
        if (ix >= 0 && ix < in.W && iy >= 0 && iy < in.H) {
            const uint8_t* q = ip + ((size_t)iy * in.W + ix) * 3;
            v0 = s_lut[q[0]]; v1 = s_lut[256 + q[1]]; v2 = s_lut[512 + q[2]];
        }
        s_in[i * 3 + 0] = v0; s_in[i * 3 + 1] = v1; s_in[i * 3 + 2] = v2;
    }
    __syncthreads();
    const int tx = tid % STEM_T, ty = tid / STEM_T;
    const int ox = ox0 + tx, oy = oy0 + ty;
    if (ox >= out.W || oy >= out.H) return;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = wts.b[c];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float v = s_in[((ty * 2 + ky) * STEM_IN + tx * 2 + kx) * 3 + ci];
#pragma unroll
                for (int co = 0; co < 16; ++co) acc[co] = fmaf(v, wts.w[((ky * 3 + kx) * 3 + ci) * 16 + co], acc[co]);
            }
    float* op = out.p + (size_t)n * out.frame_stride + ((size_t)oy * out.W + ox) * out.pix_stride;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 v = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        st4(op + q * 4, v);
    }
}

// host_w: 27*16 weights [ky][kx][ci][co] followed by 16 biases, in HOST memory
void launch_stem(const U8View& in, const float* lut, const TView& out, const float* host_w, int relu, int frames,
                 cudaStream_t s) {
    dim3 grid((out.W + STEM_T - 1) / STEM_T, (out.H + STEM_T - 1) / STEM_T, frames);
    launch_pdl(stem_kernel, grid, dim3(STEM_T * STEM_T), 0, s, in, lut, out, *reinterpret_cast<const StemWeights*>(host_w), relu);
}

// ---------------------------------------------------------------------------------------------
// K4 depthwise 3x3, pad 1, stride 1|2: one thread = one output pixel x 4 channels (float4), channel
// quads fastest so a warp reads/writes contiguous NHWC bytes; the 3x3 neighbourhood re-reads hit L1.
// ---------------------------------------------------------------------------------------------
// thread = 4 channels x one column x a strip of R vertically adjacent outputs: the 9 weights (per channel, so
// they cannot be warp-uniform constants) are loaded once, all (R-1)*S+3 input rows are fetched up front
// (independent 16-byte loads in flight) and reused across the strip, halving the loads per output.
// Grid: x = column tile, y = strip, z = frame; thread = (4-channel group, column) with the channel group fastest, so a
// warp's 16-byte accesses are contiguous along C. No index division beyond one 32-bit tid / C4.
template <int S, int R>
__global__ void __launch_bounds__(128)
depthwise3x3_kernel(TView in, TView out, const float* __restrict__ w, const float* __restrict__ b, int relu, int xt) {
    constexpr int NR = (R - 1) * S + 3;
    pdl_launch_dependents();
    pdl_wait();
    const int C4 = out.C >> 2;
    const int xl = (int)threadIdx.x / C4;
    const int c = ((int)threadIdx.x - xl * C4) * 4;
    const int x = (int)blockIdx.x * xt + xl;
    if (xl >= xt || x >= out.W) return;
    const int n = blockIdx.z;
    const int oy0 = (int)blockIdx.y * R;
    const int iy0 = oy0 * S - 1, ix0 = x * S - 1;
    float4 wv[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wv[k] = ldg4(w + k * out.C + c);
    const float4 bias = ldg4(b + c);
    const float* ip = in.p + (size_t)n * in.frame_stride + c;
    const bool cok[3] = {ix0 >= 0, true, ix0 + 2 < in.W};
    float4 v[NR][3];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int iy = iy0 + r;
        const bool rok = iy >= 0 && iy < in.H;
        const float* rp = ip + ((long long)iy * in.W + ix0) * in.pix_stride;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
            v[r][kx] = (rok && cok[kx]) ? ld4(rp + kx * in.pix_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* op = out.p + (size_t)n * out.frame_stride + ((size_t)oy0 * out.W + x) * out.pix_stride + c;
#pragma unroll
    for (int o = 0; o < R; ++o) {
        if (oy0 + o >= out.H) break;
        float4 acc = bias;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 a = v[o * S + ky][kx];
                const float4 ww = wv[ky * 3 + kx];
                acc.x = fmaf(a.x, ww.x, acc.x); acc.y = fmaf(a.y, ww.y, acc.y);
                acc.z = fmaf(a.z, ww.z, acc.z); acc.w = fmaf(a.w, ww.w, acc.w);
            }
        if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        st4(op + (size_t)o * out.W * out.pix_stride, acc);
    }
}

void launch_depthwise(const TView& in, const TView& out, const float* w_tc, const float* b, int stride, int relu,
                      int frames, cudaStream_t s) {
    const int C4 = out.C / 4;
    const int xt = 128 / C4;  // columns per CTA (the engine only routes C <= 512 here)
    if (stride == 1) {
        dim3 grid((out.W + xt - 1) / xt, (out.H + 3) / 4, frames);
        launch_pdl(depthwise3x3_kernel<1, 4>, grid, dim3(128), 0, s, in, out, w_tc, b, relu, xt);
    } else {
        dim3 grid((out.W + xt - 1) / xt, (out.H + 1) / 2, frames);
        launch_pdl(depthwise3x3_kernel<2, 2>, grid, dim3(128), 0, s, in, out, w_tc, b, relu, xt);
    }
}

// ---------------------------------------------------------------------------------------------
// K5 pointwise 1x1 = GEMM  out[M=frames*H*W][N=Cout] = in[M][K=Cin] * W[K][N] (+bias, +res, ReLU).
// SIMT fp32 (1e-4 parity contract). CTA tile BM x BN, BK = 32; A tile kept [m][k] (k contiguous,
// +4 pad), B tile [k][n]; thread (tx,ty) owns columns tx*4..+3 and the interleaved rows
// ty + i*RT, so LDS.128 of A rows from neighbouring ty hit different banks and every
// k4-step is TM + 4 LDS.128 for 16*TM FFMA.
// ---------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(256)
pointwise_kernel(TView in, TView out, TView res, int has_res, const float* __restrict__ w,
                 const float* __restrict__ b, int relu, int M, int K, int N) {
    constexpr int BK = 32, TN = 4;
    constexpr int TXN = BN / TN;        // threads along n
    constexpr int RT = 256 / TXN;       // threads along m
    constexpr int TM = BM / RT;         // rows per thread
    constexpr int LDA = BK + 4;
    __shared__ __align__(16) float As[BM * LDA];
    __shared__ __align__(16) float Bs[BK * BN];
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int HW = in.H * in.W;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: BM rows x BK floats, float4 along k (Cin % 4 == 0 guaranteed by the dispatcher)
        for (int i = tid; i < BM * (BK / 4); i += 256) {
            const int r = i / (BK / 4), kq = (i % (BK / 4)) * 4;
            const int m = m0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && k0 + kq < K) {
                const int f = m / HW;
                const int pix = m - f * HW;
                v = ld4(in.p + (size_t)f * in.frame_stride + (size_t)pix * in.pix_stride + k0 + kq);
            }
            st4(As + r * LDA + kq, v);
        }
        // B tile: BK x BN from W[K][N]
        for (int i = tid; i < BK * BN; i += 256) {
            const int kk = i / BN, nn = i % BN;
            Bs[i] = (k0 + kk < K && n0 + nn < N) ? __ldg(w + (size_t)(k0 + kk) * N + n0 + nn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kq = 0; kq < BK; kq += 4) {
            float4 bv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = ld4(Bs + (kq + j) * BN + tx * TN);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float4 a = ld4(As + (ty + i * RT) * LDA + kq);
                acc[i][0] = fmaf(a.x, bv[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, bv[0].y, acc[i][1]);
                acc[i][2] = fmaf(a.x, bv[0].z, acc[i][2]); acc[i][3] = fmaf(a.x, bv[0].w, acc[i][3]);
                acc[i][0] = fmaf(a.y, bv[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, bv[1].y, acc[i][1]);
                acc[i][2] = fmaf(a.y, bv[1].z, acc[i][2]); acc[i][3] = fmaf(a.y, bv[1].w, acc[i][3]);
                acc[i][0] = fmaf(a.z, bv[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, bv[2].y, acc[i][1]);
                acc[i][2] = fmaf(a.z, bv[2].z, acc[i][2]); acc[i][3] = fmaf(a.z, bv[2].w, acc[i][3]);
                acc[i][0] = fmaf(a.w, bv[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, bv[3].y, acc[i][1]);
                acc[i][2] = fmaf(a.w, bv[3].z, acc[i][2]); acc[i][3] = fmaf(a.w, bv[3].w, acc[i][3]);
            }
        }
        __syncthreads();
    }
    const int nb = n0 + tx * TN;
    if (nb >= N) return;
    float bias[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bias[j] = (nb + j < N) ? b[nb + j] : 0.f;
    const bool vec = (nb + TN <= N) && ((out.pix_stride & 3) == 0) && ((N & 3) == 0) &&
                     ((reinterpret_cast<size_t>(out.p) & 15) == 0) && ((out.frame_stride & 3) == 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty + i * RT;
        if (m >= M) continue;
        const int f = m / HW;
        const int pix = m - f * HW;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) v[j] = acc[i][j] + bias[j];
        if (has_res) {
            const float* rp = res.p + (size_t)f * res.frame_stride + (size_t)pix * res.pix_stride + nb;
#pragma unroll
            for (int j = 0; j < TN; ++j) if (nb + j < N) v[j] += rp[j];
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < TN; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        float* op = out.p + (size_t)f * out.frame_stride + (size_t)pix * out.pix_stride + nb;
        if (vec) {
            st4(op, make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j) if (nb + j < N) op[j] = v[j];
        }
    }
}

void launch_pointwise(const TView& in, const TView& out, const TView* res, const float* w_io, const float* b,
                      int relu, int frames, cudaStream_t s) {
    const int M = frames * in.H * in.W;
    const int K = in.C, N = out.C;
    TView r = res ? *res : TView{};
    if (N > 32) {
        dim3 grid((unsigned)((M + 127) / 128), (N + 63) / 64);
        pointwise_kernel<128, 64><<<grid, 256, 0, s>>>(in, out, r, res != nullptr, w_io, b, relu, M, K, N);
    } else if (N > 16) {
        dim3 grid((unsigned)((M + 127) / 128), 1);
        pointwise_kernel<128, 32><<<grid, 256, 0, s>>>(in, out, r, res != nullptr, w_io, b, relu, M, K, N);
    } else {
        dim3 grid((unsigned)((M + 127) / 128), 1);
        pointwise_kernel<128, 16><<<grid, 256, 0, s>>>(in, out, r, res != nullptr, w_io, b, relu, M, K, N);
    }
}

// ---------------------------------------------------------------------------------------------
// K4+K5 fused: depthwise 3x3 (+bias+ReLU) -> pointwise 1x1 (+bias, +-ReLU). Every depthwise conv
// of UltraFace feeds exactly one 1x1 conv (conv_dw blocks, the SSD `sep` heads, the extras), so
// the depthwise result never has to reach HBM: a CTA computes the depthwise output of BM
// consecutive output pixels (linear index over frames*H*W) for all C channels into shared memory
// (phase 1, float4 over channels, coalesced NHWC reads, halo re-reads served by L1) and then runs
// the BM x C x N GEMM against the 1x1 weights from there (phase 2, same register tiling as
// pointwise_kernel). gridDim.y splits N so tiny maps still fill 148 SMs (phase 1 is recomputed per
// split: 9 MAC/elt against N MAC/elt).
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
fused_dwpw_kernel(TView in, TView out, const float* __restrict__ dw_w, const float* __restrict__ dw_b, int stride,
                  int dw_relu, const float* __restrict__ pw_w, const float* __restrict__ pw_b, int pw_relu, int M,
                  int n_per_cta) {
    constexpr int TN = 4, TXN = BN / TN, RT = 256 / TXN, TM = BM / RT;
    extern __shared__ __align__(16) float smem[];
    const int C = in.C, N = out.C, LDA = C + 4;
    float* As = smem;                       // [BM][LDA] depthwise output
    float* Bs = As + BM * LDA;              // [BK][BN]
    int* s_xy = reinterpret_cast<int*>(Bs + BK * BN);  // [BM] packed (y << 16 | x), -1 = out of range
    long long* s_base = reinterpret_cast<long long*>(s_xy + BM);  // [BM] frame offset into `in`
    long long* s_obase = s_base + BM;                             // [BM] pixel offset into `out`
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int HWo = out.H * out.W;
    for (int r = tid; r < BM; r += 256) {
        const int m = m0 + r;
        if (m < M) {
            const int f = m / HWo;
            const int pix = m - f * HWo;
            const int y = pix / out.W;
            s_xy[r] = (y << 16) | (pix - y * out.W);
            s_base[r] = (long long)f * in.frame_stride;
            s_obase[r] = (long long)f * out.frame_stride + (long long)pix * out.pix_stride;
        } else {
            s_xy[r] = -1;
            s_base[r] = 0;
            s_obase[r] = 0;
        }
    }
    __syncthreads();
    // ---- phase 1: depthwise into shared memory
    const int C4 = C >> 2;
    for (int it = tid; it < BM * C4; it += 256) {
        const int r = it / C4, c = (it - r * C4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int xy = s_xy[r];
        if (xy >= 0) {
            const int y = xy >> 16, x = xy & 0xffff;
            acc = ldg4(dw_b + c);
            const float* ip = in.p + s_base[r] + c;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = y * stride - 1 + ky;
                if (iy < 0 || iy >= in.H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = x * stride - 1 + kx;
                    if (ix < 0 || ix >= in.W) continue;
                    const float4 v = ld4(ip + ((size_t)iy * in.W + ix) * in.pix_stride);
                    const float4 ww = ldg4(dw_w + (ky * 3 + kx) * C + c);
                    acc.x = fmaf(v.x, ww.x, acc.x); acc.y = fmaf(v.y, ww.y, acc.y);
                    acc.z = fmaf(v.z, ww.z, acc.z); acc.w = fmaf(v.w, ww.w, acc.w);
                }
            }
            if (dw_relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        }
        st4(As + r * LDA + c, acc);
    }
    // ---- phase 2: GEMM over this CTA's slice of N
    const int tx = tid % TXN, ty = tid / TXN;
    const int n_begin = blockIdx.y * n_per_cta;
    const int n_end = min(N, n_begin + n_per_cta);
    for (int n0 = n_begin; n0 < n_end; n0 += BN) {
        float acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < C; k0 += BK) {
            __syncthreads();  // As complete (first pass) / previous Bs consumed
            for (int i = tid; i < BK * BN; i += 256) {
                const int kk = i / BN, nn = i % BN;
                Bs[i] = (k0 + kk < C && n0 + nn < n_end) ? __ldg(pw_w + (size_t)(k0 + kk) * N + n0 + nn) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kq = 0; kq < BK; kq += 4) {
                float4 bv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = ld4(Bs + (kq + j) * BN + tx * TN);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const float4 a = ld4(As + (ty + i * RT) * LDA + k0 + kq);
                    acc[i][0] = fmaf(a.x, bv[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, bv[0].y, acc[i][1]);
                    acc[i][2] = fmaf(a.x, bv[0].z, acc[i][2]); acc[i][3] = fmaf(a.x, bv[0].w, acc[i][3]);
                    acc[i][0] = fmaf(a.y, bv[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, bv[1].y, acc[i][1]);
                    acc[i][2] = fmaf(a.y, bv[1].z, acc[i][2]); acc[i][3] = fmaf(a.y, bv[1].w, acc[i][3]);
                    acc[i][0] = fmaf(a.z, bv[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, bv[2].y, acc[i][1]);
                    acc[i][2] = fmaf(a.z, bv[2].z, acc[i][2]); acc[i][3] = fmaf(a.z, bv[2].w, acc[i][3]);
                    acc[i][0] = fmaf(a.w, bv[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, bv[3].y, acc[i][1]);
                    acc[i][2] = fmaf(a.w, bv[3].z, acc[i][2]); acc[i][3] = fmaf(a.w, bv[3].w, acc[i][3]);
                }
            }
        }
        const int nb = n0 + tx * TN;
        if (nb < n_end) {
            float bias[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) bias[j] = (nb + j < N) ? pw_b[nb + j] : 0.f;
            const bool vec = (nb + TN <= N) && ((out.pix_stride & 3) == 0) && ((N & 3) == 0) &&
                             ((reinterpret_cast<size_t>(out.p) & 15) == 0) && ((out.frame_stride & 3) == 0);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int rr = ty + i * RT;
                if (m0 + rr >= M) continue;
                float v[TN];
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    v[j] = acc[i][j] + bias[j];
                    if (pw_relu) v[j] = fmaxf(v[j], 0.f);
                }
                float* op = out.p + s_obase[rr] + nb;
                if (vec) {
                    st4(op, make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < TN; ++j) if (nb + j < N) op[j] = v[j];
                }
            }
        }
    }
}

bool fused_dwpw_supported(int C, int N) { return C % 4 == 0 && C >= 16 && C <= 256 && N >= 1; }

template <int BM, int BN, int BK>
static void launch_fused_t(const TView& in, const TView& out, const float* dw_w, const float* dw_b, int stride,
                           int dw_relu, const float* pw_w, const float* pw_b, int pw_relu, int frames, cudaStream_t s) {
    const int M = frames * out.H * out.W;
    const int N = out.C, C = in.C;
    const int m_tiles = (M + BM - 1) / BM;
    const int n_tiles = (N + BN - 1) / BN;
    // split N across CTAs only while the grid is smaller than ~2 waves
    int split = 1;
    while (split < n_tiles && (long long)m_tiles * split < 2 * 148) ++split;
    const int tiles_per = (n_tiles + split - 1) / split;
    const int n_per_cta = tiles_per * BN;
    const int gy = (N + n_per_cta - 1) / n_per_cta;
    const size_t smem = (size_t)(BM * (C + 4) + BK * BN) * sizeof(float) + BM * (sizeof(int) + 2 * sizeof(long long));
    auto kern = fused_dwpw_kernel<BM, BN, BK>;
    static bool configured[64] = {};  // per template instantiation and device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        configured[dev & 63] = true;
    }
    kern<<<dim3(m_tiles, gy), 256, smem, s>>>(in, out, dw_w, dw_b, stride, dw_relu, pw_w, pw_b, pw_relu, M, n_per_cta);
}

void launch_fused_dwpw(const TView& in, const TView& out, const float* dw_w_tc, const float* dw_b, int stride,
                       int dw_relu, const float* pw_w_io, const float* pw_b, int pw_relu, int frames, cudaStream_t s) {
    const int C = in.C, N = out.C;
#define UF_F(BM, BN, BK) launch_fused_t<BM, BN, BK>(in, out, dw_w_tc, dw_b, stride, dw_relu, pw_w_io, pw_b, pw_relu, frames, s)
    if (C == 16) {
        if (N > 16) UF_F(128, 32, 16); else UF_F(128, 16, 16);
    } else if (C > 128) {
        if (N > 32) UF_F(64, 64, 32); else if (N > 16) UF_F(64, 32, 32); else UF_F(64, 16, 32);
    } else {
        if (N > 32) UF_F(128, 64, 32); else if (N > 16) UF_F(128, 32, 32); else UF_F(128, 16, 32);
    }
#undef UF_F
}

// ---------------------------------------------------------------------------------------------
// K6b: 3x3 conv with many input channels and few outputs (the last SSD heads: 256 -> 6 / 12 on the
// 4x5 (8x10) map). One warp per output pixel: lanes split Cin in float4 chunks (coalesced 512 B
// reads of the NHWC pixel and of the [tap][ci][co] weights), N accumulators per lane, warp-shuffle
// reduction at the end. K = 9*Cin = 2304 is far too deep for one thread per output.
// ---------------------------------------------------------------------------------------------
// CTA = P horizontally adjacent outputs, warp = one of the 9 taps: the N float4 weight loads of a (tap, channel
// chunk) are shared by the P pixels, each warp's dependent-load chain is Cin/128 trips instead of 9*Cin/128, and the
// nine partial sums meet in shared memory. (One pixel per warp re-read all 9*Cin*N weights for every pixel and sat
// at 90 % L1 throughput, 12 % issue utilisation.)
template <int N, int P>
__global__ void __launch_bounds__(288)
conv3x3_warp_kernel(TView in, TView out, const float* __restrict__ w, const float* __restrict__ b, int dil, int relu,
                    int segs) {
    __shared__ float part[9][P * N];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, tap = threadIdx.x >> 5;
    const int seg = blockIdx.x % segs;
    const int t = blockIdx.x / segs;
    const int y = t % out.H, f = t / out.H;
    const int x0 = seg * P;
    const int C = in.C;
    const int ky = tap / 3, kx = tap - ky * 3;
    float acc[P][N];
#pragma unroll
    for (int q = 0; q < P; ++q)
#pragma unroll
        for (int j = 0; j < N; ++j) acc[q][j] = 0.f;
    const int iy = y + (ky - 1) * dil;
    if (iy >= 0 && iy < in.H) {
        const float* wt = w + (size_t)tap * C * N;
        const float* row = in.p + (size_t)f * in.frame_stride + (size_t)iy * in.W * in.pix_stride;
        for (int c = lane * 4; c < C; c += 128) {
            float4 v[P];
#pragma unroll
            for (int q = 0; q < P; ++q) {
                const int ix = x0 + q + (kx - 1) * dil;
                v[q] = (ix >= 0 && ix < in.W) ? ld4(row + (size_t)ix * in.pix_stride + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float4* wp = reinterpret_cast<const float4*>(wt + (size_t)c * N);  // 4*N contiguous floats
            float wv[4 * N];
#pragma unroll
            for (int q = 0; q < N; ++q) {
                const float4 t4 = __ldg(wp + q);
                wv[q * 4] = t4.x; wv[q * 4 + 1] = t4.y; wv[q * 4 + 2] = t4.z; wv[q * 4 + 3] = t4.w;
            }
#pragma unroll
            for (int q = 0; q < P; ++q)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    acc[q][j] = fmaf(v[q].x, wv[j], acc[q][j]);
                    acc[q][j] = fmaf(v[q].y, wv[N + j], acc[q][j]);
                    acc[q][j] = fmaf(v[q].z, wv[2 * N + j], acc[q][j]);
                    acc[q][j] = fmaf(v[q].w, wv[3 * N + j], acc[q][j]);
                }
        }
    }
#pragma unroll
    for (int q = 0; q < P; ++q)
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[q][j] += __shfl_xor_sync(0xffffffffu, acc[q][j], o);
            if (lane == 0) part[tap][q * N + j] = acc[q][j];
        }
    __syncthreads();
    const int i = threadIdx.x;
    if (i < P * N) {
        const int q = i / N, j = i - q * N;
        if (x0 + q < out.W) {
            float r = b[j];
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) r += part[tp][i];
            if (relu) r = fmaxf(r, 0.f);
            out.p[(size_t)f * out.frame_stride + ((size_t)y * out.W + x0 + q) * out.pix_stride + j] = r;
        }
    }
}

bool conv3x3_warp_supported(int cin, int cout) {
    return cin % 4 == 0 && cin >= 64 && (cout == 4 || cout == 6 || cout == 8 || cout == 12 || cout == 16);
}

void launch_conv3x3_warp(const TView& in, const TView& out, const float* w_kkio, const float* b, int dil, int relu,
                         int frames, cudaStream_t s) {
    // P = 5 pixels per warp when the rows split evenly into fives (the 5x4 / 10x8 last maps), else 4
    const bool five = out.W % 5 == 0;
    const int P = five ? 5 : 4;
    const int segs = (out.W + P - 1) / P;
    const int grid = frames * out.H * segs;
#define UF_W(NN)                                                                                                                \
    do {                                                                                                                        \
        if (five) launch_pdl(conv3x3_warp_kernel<NN, 5>, dim3(grid), dim3(288), 0, s, in, out, w_kkio, b, dil, relu, segs); \
        else launch_pdl(conv3x3_warp_kernel<NN, 4>, dim3(grid), dim3(288), 0, s, in, out, w_kkio, b, dil, relu, segs);      \
    } while (0)
    switch (out.C) {
        case 4: UF_W(4); break;
        case 6: UF_W(6); break;
        case 8: UF_W(8); break;
        case 12: UF_W(12); break;
        case 16: UF_W(16); break;
    }
#undef UF_W
}

// Stage an IH x IW pixel tile (C4 float4 channel groups per pixel, top-left map coordinate (gy0, gx0)) into shared
// memory at a pixel pitch of P floats, zeros outside the map. U independent loads are issued before the first
// store: with one load per trip the loop is a chain of L2 latencies (measured 11 us per CTA on the 128-channel
// heads), which is what bounds these small-tile kernels.
template <int U>
__device__ __forceinline__ void stage_tile(const float* __restrict__ ip, const TView& in, float* __restrict__ s_in, int gy0,
                                           int gx0, int IH, int IW, int C4, int P, int tid, int nthr) {
    const int total = IH * IW * C4;
    for (int i0 = tid; i0 < total; i0 += U * nthr) {
        float4 v[U];
        int so[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * nthr;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            so[u] = -1;
            if (i < total) {
                const int pix = i / C4, q = i - pix * C4;
                const int py = pix / IW, px = pix - py * IW;
                const int gy = gy0 + py, gx = gx0 + px;
                so[u] = pix * P + q * 4;
                if (gy >= 0 && gy < in.H && gx >= 0 && gx < in.W) v[u] = ld4(ip + ((size_t)gy * in.W + gx) * in.pix_stride + q * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (so[u] >= 0) st4(s_in + so[u], v[u]);
    }
}

// ---------------------------------------------------------------------------------------------
// K4+K5 fused, pixel-per-thread form for C <= 64: CTA = 8 x TY output pixels of one frame. The
// input tile (+1 halo, stride-scaled) is staged in shared memory with coalesced float4 loads and a
// padded pixel pitch of C+4 floats (conflict-free LDS.128 for 8 neighbouring pixels); each thread
// keeps the C depthwise results of its pixel in registers and runs the 1x1 conv against weights
// broadcast from shared memory (one LDS.128 per 4 FFMA, no barriers after the fill), NT outputs
// at a time. Used for the large, memory-bound maps (16/32/64 channels).
// ---------------------------------------------------------------------------------------------
template <int C, int S, int TY, int NT>
__global__ void __launch_bounds__(8 * TY)
fused_dwpw_pix_kernel(TView in, TView out, const float* __restrict__ dw_w, const float* __restrict__ dw_b, int dw_relu,
                      const float* __restrict__ pw_w, const float* __restrict__ pw_b, int pw_relu, int tiles_x,
                      int tiles_y, int n_pad) {
    constexpr int TX = 8, IW = (TX - 1) * S + 3, IH = (TY - 1) * S + 3, P = C + 4, NTHR = TX * TY, C4 = C / 4;
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;                      // IH*IW*P
    float* s_dw = s_in + IH * IW * P;        // 9*C + C (bias)
    float* s_pw = s_dw + 10 * C;             // C * n_pad
    float* s_pb = s_pw + C * n_pad;          // n_pad
    const int tid = threadIdx.x;
    const int N = out.C;
    int bid = blockIdx.x;
    const int txi = bid % tiles_x; bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int f = bid / tiles_y;
    const int x0 = txi * TX, y0 = tyi * TY;
    const int gx0 = x0 * S - 1, gy0 = y0 * S - 1;
    pdl_launch_dependents();
    for (int i = tid; i < 9 * C; i += NTHR) s_dw[i] = dw_w[i];
    for (int i = tid; i < C; i += NTHR) s_dw[9 * C + i] = dw_b[i];
    for (int i = tid; i < C * n_pad; i += NTHR) {
        const int ci = i / n_pad, n = i - ci * n_pad;
        s_pw[i] = n < N ? pw_w[(size_t)ci * N + n] : 0.f;
    }
    for (int i = tid; i < n_pad; i += NTHR) s_pb[i] = i < N ? pw_b[i] : 0.f;
    pdl_wait();  // weights above are static; the input tile below is the predecessor's output
    const float* ip = in.p + (size_t)f * in.frame_stride;
    stage_tile<4>(ip, in, s_in, gy0, gx0, IH, IW, C4, P, tid, NTHR);
    __syncthreads();
    const int tx = tid % TX, ty = tid / TX;
    const int ox = x0 + tx, oy = y0 + ty;
    if (ox >= out.W || oy >= out.H) return;
    float d[C];
#pragma unroll
    for (int c = 0; c < C; c += 4) {
        float4 a = ld4(s_dw + 9 * C + c);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 v = ld4(s_in + ((ty * S + ky) * IW + tx * S + kx) * P + c);
                const float4 ww = ld4(s_dw + (ky * 3 + kx) * C + c);
                a.x = fmaf(v.x, ww.x, a.x); a.y = fmaf(v.y, ww.y, a.y); a.z = fmaf(v.z, ww.z, a.z); a.w = fmaf(v.w, ww.w, a.w);
            }
        if (dw_relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        d[c] = a.x; d[c + 1] = a.y; d[c + 2] = a.z; d[c + 3] = a.w;
    }
    float* op = out.p + (size_t)f * out.frame_stride + ((size_t)oy * out.W + ox) * out.pix_stride;
    const bool vec = ((N & 3) == 0) && ((out.pix_stride & 3) == 0) && ((reinterpret_cast<size_t>(out.p) & 15) == 0) &&
                     ((out.frame_stride & 3) == 0);
    for (int n0 = 0; n0 < n_pad; n0 += NT) {
        float o[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) o[j] = s_pb[n0 + j];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            const float a = d[ci];
            const float4* wp = reinterpret_cast<const float4*>(s_pw + ci * n_pad + n0);
#pragma unroll
            for (int q = 0; q < NT / 4; ++q) {
                const float4 ww = wp[q];
                o[q * 4 + 0] = fmaf(a, ww.x, o[q * 4 + 0]); o[q * 4 + 1] = fmaf(a, ww.y, o[q * 4 + 1]);
                o[q * 4 + 2] = fmaf(a, ww.z, o[q * 4 + 2]); o[q * 4 + 3] = fmaf(a, ww.w, o[q * 4 + 3]);
            }
        }
        if (pw_relu) {
#pragma unroll
            for (int j = 0; j < NT; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        if (vec && n0 + NT <= N) {
#pragma unroll
            for (int q = 0; q < NT / 4; ++q) st4(op + n0 + q * 4, make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]));
        } else {
#pragma unroll
            for (int j = 0; j < NT; ++j) if (n0 + j < N) op[n0 + j] = o[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// SSD `sep` heads (dw3x3 + ReLU -> 1x1 C -> 4..16) on the 64 / 128 / 256-channel maps: same tile scheme as above, but
// the depthwise and pointwise weights are kernel-parameter constants (6.8 - 27 KB), so the arithmetic is pure FFMA
// with constant-bank operands and shared memory only serves the input tile. For C > 64 the channels are split over
// SPLIT thread groups (each with compile-time channel indices, hence the PART template recursion) whose partial
// sums meet in shared memory: the wide heads sit on tiny maps, and a pixel-per-thread CTA alone would leave the SM
// almost empty. Replaces depthwise kernel + tensor-core GEMM (2 launches, ~36 us) for those layers.
// ---------------------------------------------------------------------------------------------
template <int C, int NP>
struct HeadWeights {
    float dw[9 * C];
    float dwb[C];
    float pw[C * NP];  // [ci][n], zero padded to NP outputs
    float pwb[NP];
};

template <int C, int NP, int IW, int SPLIT, int PART>
__device__ __forceinline__ void head_part(const float* __restrict__ s_px, const HeadWeights<C, NP>& wts, int dw_relu, float (&o)[NP]) {
    constexpr int P = C + 4, C0 = PART * (C / SPLIT), C1 = C0 + C / SPLIT;
#pragma unroll
    for (int c = C0; c < C1; c += 4) {
        float a0 = wts.dwb[c], a1 = wts.dwb[c + 1], a2 = wts.dwb[c + 2], a3 = wts.dwb[c + 3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 v = ld4(s_px + (ky * IW + kx) * P + c);
                const int wo = (ky * 3 + kx) * C + c;
                a0 = fmaf(v.x, wts.dw[wo], a0); a1 = fmaf(v.y, wts.dw[wo + 1], a1);
                a2 = fmaf(v.z, wts.dw[wo + 2], a2); a3 = fmaf(v.w, wts.dw[wo + 3], a3);
            }
        if (dw_relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
#pragma unroll
        for (int j = 0; j < NP; ++j) {  // the depthwise result is consumed at once: no C-long register array
            o[j] = fmaf(a0, wts.pw[c * NP + j], o[j]);
            o[j] = fmaf(a1, wts.pw[(c + 1) * NP + j], o[j]);
            o[j] = fmaf(a2, wts.pw[(c + 2) * NP + j], o[j]);
            o[j] = fmaf(a3, wts.pw[(c + 3) * NP + j], o[j]);
        }
    }
}

// one head on the staged tile: every thread of the CTA must call this (it contains a barrier when SPLIT > 1)
template <int C, int NP, int TX, int TY, int SPLIT>
__device__ __forceinline__ void head_compute(const float* __restrict__ s_in, float* __restrict__ s_part,
                                             const HeadWeights<C, NP>& wts, int dw_relu, int pw_relu, const TView& out, int f,
                                             int x0, int y0) {
    constexpr int IW = TX + 2, P = C + 4, GT = (TX * TY + 31) / 32 * 32;
    const int tid = threadIdx.x;
    const int part = tid / GT, lt = tid - part * GT;  // part is warp-uniform: GT is a multiple of 32
    const int tx = lt % TX, ty = lt / TX;
    const int ox = x0 + tx, oy = y0 + ty;
    const bool active = lt < TX * TY && ox < out.W && oy < out.H;
    float o[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) o[j] = 0.f;
    if (active) {
        const float* s_px = s_in + (ty * IW + tx) * P;
        if (part == 0) {
#pragma unroll
            for (int j = 0; j < NP; ++j) o[j] = wts.pwb[j];
            head_part<C, NP, IW, SPLIT, 0>(s_px, wts, dw_relu, o);
        }
        if (SPLIT > 1 && part == 1) head_part<C, NP, IW, SPLIT, (SPLIT > 1 ? 1 : 0)>(s_px, wts, dw_relu, o);
        if (SPLIT > 2 && part == 2) head_part<C, NP, IW, SPLIT, (SPLIT > 2 ? 2 : 0)>(s_px, wts, dw_relu, o);
        if (SPLIT > 3 && part == 3) head_part<C, NP, IW, SPLIT, (SPLIT > 3 ? 3 : 0)>(s_px, wts, dw_relu, o);
    }
    if (SPLIT > 1) {  // s_part: [SPLIT-1][TX*TY][NP]
        if (active && part > 0) {
#pragma unroll
            for (int j = 0; j < NP; ++j) s_part[((part - 1) * TX * TY + lt) * NP + j] = o[j];
        }
        __syncthreads();
        if (active && part == 0) {
#pragma unroll
            for (int sp = 0; sp < SPLIT - 1; ++sp)
#pragma unroll
                for (int j = 0; j < NP; ++j) o[j] += s_part[(sp * TX * TY + lt) * NP + j];
        }
    }
    if (!active || part != 0) return;
    float* op = out.p + (size_t)f * out.frame_stride + ((size_t)oy * out.W + ox) * out.pix_stride;
    const int N = out.C;
#pragma unroll
    for (int j = 0; j < NP; ++j)
        if (j < N) op[j] = pw_relu ? fmaxf(o[j], 0.f) : o[j];
}

template <int C, int NP, int TX, int TY, int SPLIT>
__global__ void __launch_bounds__(((TX * TY + 31) / 32 * 32) * SPLIT)
head_dwpw_kernel(TView in, TView out, const __grid_constant__ HeadWeights<C, NP> wts, int dw_relu, int pw_relu, int tiles_x,
                 int tiles_y) {
    constexpr int IW = TX + 2, IH = TY + 2, P = C + 4, NTHR = ((TX * TY + 31) / 32 * 32) * SPLIT, C4 = C / 4;
    extern __shared__ __align__(16) float s_in[];  // IH*IW*P, then (SPLIT-1) * TX*TY * NP partial sums
    pdl_launch_dependents();
    pdl_wait();
    int bid = blockIdx.x;
    const int txi = bid % tiles_x; bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int f = bid / tiles_y;
    const int x0 = txi * TX, y0 = tyi * TY;
    stage_tile<8>(in.p + (size_t)f * in.frame_stride, in, s_in, y0 - 1, x0 - 1, IH, IW, C4, P, threadIdx.x, NTHR);
    __syncthreads();
    head_compute<C, NP, TX, TY, SPLIT>(s_in, s_in + IH * IW * P, wts, dw_relu, pw_relu, out, f, x0, y0);
}

// The class and box heads of one map read the same input: one staged tile serves both (NPA = 8, NPB = 16 outputs).
template <int C>
struct Head2Weights {
    HeadWeights<C, 8> a;
    HeadWeights<C, 16> b;
};

template <int C, int TX, int TY, int SPLIT>
__global__ void __launch_bounds__(((TX * TY + 31) / 32 * 32) * SPLIT)
head2_dwpw_kernel(TView in, TView out_a, TView out_b, const __grid_constant__ Head2Weights<C> wts, int relu_bits, int tiles_x,
                  int tiles_y) {
    constexpr int IW = TX + 2, IH = TY + 2, P = C + 4, NTHR = ((TX * TY + 31) / 32 * 32) * SPLIT, C4 = C / 4;
    extern __shared__ __align__(16) float s_in[];  // IH*IW*P, then (SPLIT-1) * TX*TY * 16 partial sums
    pdl_launch_dependents();
    pdl_wait();
    int bid = blockIdx.x;
    const int txi = bid % tiles_x; bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int f = bid / tiles_y;
    const int x0 = txi * TX, y0 = tyi * TY;
    stage_tile<8>(in.p + (size_t)f * in.frame_stride, in, s_in, y0 - 1, x0 - 1, IH, IW, C4, P, threadIdx.x, NTHR);
    __syncthreads();
    float* s_part = s_in + IH * IW * P;
    head_compute<C, 8, TX, TY, SPLIT>(s_in, s_part, wts.a, relu_bits & 1, relu_bits & 2, out_a, f, x0, y0);
    if (SPLIT > 1) __syncthreads();  // the partial-sum buffer is reused
    head_compute<C, 16, TX, TY, SPLIT>(s_in, s_part, wts.b, relu_bits & 4, relu_bits & 8, out_b, f, x0, y0);
}

bool head_dwpw_supported(int C, int N, int stride) { return (C == 64 || C == 128 || C == 256) && stride == 1 && N >= 1 && N <= 16; }
size_t head_dwpw_weight_floats(int C, int N) { const int np = N <= 8 ? 8 : 16; return 10 * (size_t)C + (size_t)C * np + np; }

template <int C, int NP, int TX, int TY, int SPLIT>
static void launch_head_t(const TView& in, const TView& out, const float* host_w, int dw_relu, int pw_relu, int frames,
                          cudaStream_t s) {
    static_assert(sizeof(HeadWeights<C, NP>) + 2 * sizeof(TView) + 64 <= 32764, "kernel parameter space exceeded");
    constexpr int GT = (TX * TY + 31) / 32 * 32;
    const int tiles_x = (out.W + TX - 1) / TX, tiles_y = (out.H + TY - 1) / TY;
    const size_t smem = ((size_t)(TY + 2) * (TX + 2) * (C + 4) + (size_t)(SPLIT - 1) * TX * TY * NP) * sizeof(float);
    auto kern = head_dwpw_kernel<C, NP, TX, TY, SPLIT>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        configured[dev & 63] = true;
    }
    launch_pdl(kern, dim3(tiles_x * tiles_y * frames), dim3(GT * SPLIT), smem, s, in, out,
               *reinterpret_cast<const HeadWeights<C, NP>*>(host_w), dw_relu, pw_relu, tiles_x, tiles_y);
}

// host_w: [dw 9*C][dw bias C][pw C*NP (ci major, zero padded)][pw bias NP] with NP = 8 (N <= 8) or 16, HOST memory
void launch_head_dwpw(const TView& in, const TView& out, const float* host_w, int dw_relu, int pw_relu, int frames,
                      cudaStream_t s) {
#define UF_H(CC, TXX, TYY, SS)                                                                          \
    do {                                                                                                \
        if (out.C <= 8) launch_head_t<CC, 8, TXX, TYY, SS>(in, out, host_w, dw_relu, pw_relu, frames, s); \
        else launch_head_t<CC, 16, TXX, TYY, SS>(in, out, host_w, dw_relu, pw_relu, frames, s);           \
    } while (0)
    // (smaller tiles / more channel slices for occupancy were measured: no gain - ncu shows these kernels limited by
    // the constant-load stream, one LDCU.128 per 4 FFMA, not by latency hiding)
    if (in.C == 64) UF_H(64, 8, 32, 1);
    else if (in.C == 128) UF_H(128, 20, 5, 2);   // 20x15 map: three tiles of 100 pixels, two channel halves
    else if (in.C == 256) UF_H(256, 10, 8, 4);   // 10x8 map: one tile, four channel quarters
#undef UF_H
}

bool head2_dwpw_supported(int C, int Na, int Nb) { return (C == 64 || C == 128) && Na >= 1 && Na <= 8 && Nb >= 1 && Nb <= 16; }

template <int C, int TX, int TY, int SPLIT>
static void launch_head2_t(const TView& in, const TView& out_a, const TView& out_b, const float* host_w, int relu_bits, int frames,
                           cudaStream_t s) {
    static_assert(sizeof(Head2Weights<C>) + 3 * sizeof(TView) + 64 <= 32764, "kernel parameter space exceeded");
    constexpr int GT = (TX * TY + 31) / 32 * 32;
    const int tiles_x = (out_a.W + TX - 1) / TX, tiles_y = (out_a.H + TY - 1) / TY;
    const size_t smem = ((size_t)(TY + 2) * (TX + 2) * (C + 4) + (size_t)(SPLIT - 1) * TX * TY * 16) * sizeof(float);
    auto kern = head2_dwpw_kernel<C, TX, TY, SPLIT>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        configured[dev & 63] = true;
    }
    launch_pdl(kern, dim3(tiles_x * tiles_y * frames), dim3(GT * SPLIT), smem, s, in, out_a, out_b,
               *reinterpret_cast<const Head2Weights<C>*>(host_w), relu_bits, tiles_x, tiles_y);
}

// host_w: HeadWeights<C, 8> of head A followed by HeadWeights<C, 16> of head B (HOST memory);
// relu_bits: 1 = A depthwise, 2 = A pointwise, 4 = B depthwise, 8 = B pointwise
void launch_head2_dwpw(const TView& in, const TView& out_a, const TView& out_b, const float* host_w, int relu_bits, int frames,
                       cudaStream_t s) {
    if (in.C == 64) launch_head2_t<64, 8, 32, 1>(in, out_a, out_b, host_w, relu_bits, frames, s);
    else if (in.C == 128) launch_head2_t<128, 20, 5, 2>(in, out_a, out_b, host_w, relu_bits, frames, s);
}

bool fused_dwpw_pix_supported(int C, int N, int stride) {
    if (N > 64) return false;
    if (C == 16 || C == 32) return stride == 1 || stride == 2;
    if (C == 64) return stride == 1;
    return false;
}

template <int C, int S, int TY, int NT>
static void launch_pix_t(const TView& in, const TView& out, const float* dw_w, const float* dw_b, int dw_relu,
                         const float* pw_w, const float* pw_b, int pw_relu, int frames, cudaStream_t s) {
    constexpr int TX = 8, IW = (TX - 1) * S + 3, IH = (TY - 1) * S + 3;
    const int n_pad = (out.C + NT - 1) / NT * NT;
    const int tiles_x = (out.W + TX - 1) / TX, tiles_y = (out.H + TY - 1) / TY;
    const size_t smem = (size_t)(IH * IW * (C + 4) + 10 * C + C * n_pad + n_pad) * sizeof(float);
    auto kern = fused_dwpw_pix_kernel<C, S, TY, NT>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured[dev & 63] = true;
    }
    launch_pdl(kern, dim3(tiles_x * tiles_y * frames), dim3(TX * TY), smem, s, in, out, dw_w, dw_b, dw_relu, pw_w, pw_b, pw_relu,
               tiles_x, tiles_y, n_pad);
}

void launch_fused_dwpw_pix(const TView& in, const TView& out, const float* dw_w_tc, const float* dw_b, int stride,
                           int dw_relu, const float* pw_w_io, const float* pw_b, int pw_relu, int frames,
                           cudaStream_t s) {
    const int C = in.C, N = out.C;
#define UF_P(CC, SS, TYY, NTT) launch_pix_t<CC, SS, TYY, NTT>(in, out, dw_w_tc, dw_b, dw_relu, pw_w_io, pw_b, pw_relu, frames, s)
    if (C == 16) { if (stride == 1) UF_P(16, 1, 32, 32); else UF_P(16, 2, 16, 32); }
    else if (C == 32) { if (stride == 1) UF_P(32, 1, 32, 32); else UF_P(32, 2, 16, 32); }
    else if (C == 64) { if (N > 16) UF_P(64, 1, 32, 32); else if (N > 8) UF_P(64, 1, 32, 16); else UF_P(64, 1, 32, 8); }
#undef UF_P
}

// ---------------------------------------------------------------------------------------------
// K6 small dense 3x3 (RFB branches: 8->16, 8->12, 12->16, 16->16 with dilation 1/2/3/5; stride 1,
// pad == dil). Weights [3][3][CIN][COUT] in shared memory, read as broadcast float4; each thread
// computes two horizontally adjacent pixels x COUT so one weight LDS.128 feeds 8 FFMA.
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct SmallDenseWeights {
    float w[9 * CIN * COUT];  // [ky][kx][ci][co]
    float b[COUT];
};

// CTA = 8 x 16 output pixels of one frame, thread = pixel x all COUT. The input tile (+dilation halo) is staged
// in shared memory with coalesced float4 loads and a padded pixel pitch (conflict-free LDS.128); weights are
// constant-bank operands, so the inner loop is pure FFMA: 9*CIN*COUT per thread against 9*CIN/4 LDS.128.
constexpr int SD_TX = 8, SD_TY = 16;

template <int CIN, int COUT>
__global__ void __launch_bounds__(SD_TX * SD_TY)
small_dense3x3_kernel(TView in, TView out, const __grid_constant__ SmallDenseWeights<CIN, COUT> wts, int dil, int relu,
                      int tiles_x, int tiles_y) {
    constexpr int P = CIN + 4, C4 = CIN / 4, NTHR = SD_TX * SD_TY;
    extern __shared__ __align__(16) float s_in[];  // (SD_TY+2d) x (SD_TX+2d) x P
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int txi = bid % tiles_x; bid /= tiles_x;
    const int tyi = bid % tiles_y;
    const int f = bid / tiles_y;
    const int x0 = txi * SD_TX, y0 = tyi * SD_TY;
    const int IW = SD_TX + 2 * dil, IH = SD_TY + 2 * dil;
    const float* ip = in.p + (size_t)f * in.frame_stride;
    stage_tile<4>(ip, in, s_in, y0 - dil, x0 - dil, IH, IW, C4, P, tid, NTHR);
    __syncthreads();
    const int tx = tid % SD_TX, ty = tid / SD_TX;
    const int ox = x0 + tx, oy = y0 + ty;
    if (ox >= out.W || oy >= out.H) return;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = wts.b[c];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float* sp = s_in + ((ty + ky * dil) * IW + tx + kx * dil) * P;
#pragma unroll
            for (int cq = 0; cq < CIN; cq += 4) {
                const float4 a = ld4(sp + cq);
                const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                    for (int co = 0; co < COUT; ++co)
                        acc[co] = fmaf(av[ci], wts.w[((ky * 3 + kx) * CIN + cq + ci) * COUT + co], acc[co]);
            }
        }
    float* op = out.p + (size_t)f * out.frame_stride + ((size_t)oy * out.W + ox) * out.pix_stride;
#pragma unroll
    for (int q = 0; q < COUT / 4; ++q) {
        float4 v = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        st4(op + q * 4, v);
    }
}

bool small_dense_supported(int cin, int cout) {
    return (cin == 8 && (cout == 16 || cout == 12)) || (cin == 12 && cout == 16) || (cin == 16 && cout == 16);
}

template <int CIN, int COUT>
static void launch_sd_t(const TView& in, const TView& out, const float* host_w, int dil, int relu, int frames, cudaStream_t s) {
    const int tiles_x = (out.W + SD_TX - 1) / SD_TX, tiles_y = (out.H + SD_TY - 1) / SD_TY;
    const size_t smem = (size_t)(SD_TY + 2 * dil) * (SD_TX + 2 * dil) * (CIN + 4) * sizeof(float);
    auto kern = small_dense3x3_kernel<CIN, COUT>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        configured[dev & 63] = true;
    }
    launch_pdl(kern, dim3(tiles_x * tiles_y * frames), dim3(SD_TX * SD_TY), smem, s, in, out,
               *reinterpret_cast<const SmallDenseWeights<CIN, COUT>*>(host_w), dil, relu, tiles_x, tiles_y);
}

// host_w: 9*CIN*COUT weights [ky][kx][ci][co] followed by COUT biases, in HOST memory; dil <= 8
void launch_small_dense(const TView& in, const TView& out, const float* host_w, int dil, int relu, int frames,
                        cudaStream_t s) {
    if (in.C == 8 && out.C == 16) launch_sd_t<8, 16>(in, out, host_w, dil, relu, frames, s);
    else if (in.C == 8 && out.C == 12) launch_sd_t<8, 12>(in, out, host_w, dil, relu, frames, s);
    else if (in.C == 12 && out.C == 16) launch_sd_t<12, 16>(in, out, host_w, dil, relu, frames, s);
    else if (in.C == 16 && out.C == 16) launch_sd_t<16, 16>(in, out, host_w, dil, relu, frames, s);
}

// ---------------------------------------------------------------------------------------------
// Fallback element-wise ops (only reached by graphs whose Add/Relu/Concat do not fuse) and the
// NHWC -> NCHW reader behind uf_tensor_read.
// ---------------------------------------------------------------------------------------------
__global__ void eltwise_kernel(TView a, TView b, int has_b, TView out, int relu, long long total) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % out.C);
        long long pix = idx / out.C;
        const int hw = out.H * out.W;
        const long long n = pix / hw;
        const int p = (int)(pix - n * hw);
        float v = a.p[n * a.frame_stride + (size_t)p * a.pix_stride + c];
        if (has_b) v += b.p[n * b.frame_stride + (size_t)p * b.pix_stride + c];
        if (relu) v = fmaxf(v, 0.f);
        out.p[n * out.frame_stride + (size_t)p * out.pix_stride + c] = v;
    }
}

void launch_add(const TView& a, const TView& b, const TView& out, int relu, int frames, cudaStream_t s) {
    long long total = (long long)frames * out.H * out.W * out.C;
    eltwise_kernel<<<grid_for(total, 256), 256, 0, s>>>(a, b, 1, out, relu, total);
}
void launch_relu(const TView& a, const TView& out, int frames, cudaStream_t s) {
    long long total = (long long)frames * out.H * out.W * out.C;
    eltwise_kernel<<<grid_for(total, 256), 256, 0, s>>>(a, TView{}, 0, out, 1, total);
}
void launch_copy(const TView& a, const TView& out, int frames, cudaStream_t s) {
    long long total = (long long)frames * out.H * out.W * out.C;
    eltwise_kernel<<<grid_for(total, 256), 256, 0, s>>>(a, TView{}, 0, out, 0, total);
}

__global__ void nhwc_to_nchw_kernel(TView a, int frame, float* __restrict__ out) {
    const int total = a.C * a.H * a.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i / (a.H * a.W), p = i - c * (a.H * a.W);
        out[i] = a.p[(size_t)frame * a.frame_stride + (size_t)p * a.pix_stride + c];
    }
}
void launch_nhwc_to_nchw(const TView& a, int frame, float* out, cudaStream_t s) {
    int total = a.C * a.H * a.W;
    nhwc_to_nchw_kernel<<<(total + 255) / 256, 256, 0, s>>>(a, frame, out);
}

}  // namespace uf
