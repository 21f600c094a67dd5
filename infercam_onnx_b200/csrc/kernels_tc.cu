// kernels_tc.cu — K5 on the 5th-generation tensor cores: pointwise (1x1) convolution as a GEMM
//   out[M = frames*H*W][N = Cout] = in[M][K = Cin] * W[N][K]^T (+bias, +residual, ReLU)
// with tcgen05.mma (kind::tf32), accumulators in TMEM, operands staged by TMA (128B swizzle),
// for the layers where the channel counts make it a real dense contraction (K >= 32).
//
// Accuracy: the reference computes in fp32 and the contract is 1e-4 abs on the raw outputs
// (BASELINE.json north_star). One TF32 pass (10-bit mantissa) does not meet that through 15 layers,
// so every product is split 3xTF32:  a = a_hi + a_lo,  w = w_hi + w_lo  (hi = top 19 bits, exact
// split) and  a*w ~= a_lo*w_hi + a_hi*w_lo + a_hi*w_hi  accumulated in fp32 in TMEM (the dropped
// a_lo*w_lo term is ~2^-22 relative). w_hi/w_lo are precomputed at load; a_hi/a_lo are produced in
// shared memory by four converter warps between the TMA and the MMA.
//
// Warp roles (320 threads, 1 CTA/SM, persistent over 128-row tiles):
//   warps 0-3  converters : wait full[s] -> split the 128x32 fp32 A block in place into hi / lo
//                           -> fence.proxy.async -> arrive conv[s]
//   warps 4-7  epilogue   : wait acc_full[a] -> tcgen05.ld TMEM -> bias/residual/ReLU -> global
//                           -> arrive acc_empty[a]     (warp w may touch TMEM lanes 32*(w%4)..+31)
//   warp  8    TMA producer (one lane): A block [128 rows][32 k], W_hi and W_lo blocks [N][32 k]
//   warp  9    MMA issuer (one lane) + TMEM allocation; tcgen05.commit frees smem stages / publishes
//              the accumulator. Two accumulator buffers in TMEM overlap the epilogue of tile i with
//              the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

// dense3x3_tc_kernel operand layout: 1 = pixel-major rows with the 32B / 64B swizzle (keyed on absolute shared-memory
// address bits, which is what makes a shifted descriptor start address legal), 0 = no-swizzle chunk-major planes.
// Both are parity-green and equally fast; see the kernel's header comment.
#define UF_D3_SWZ 1
#include "kernels.h"
#include "pdl.cuh"

namespace uf {

// 3xTF32 split of an fp32 operand: hi = the value with its low 13 mantissa bits cut (what kind::tf32 reads from the raw
// fp32 word, so hi is never materialised), lo = v - hi (exact, 13 significant bits) ROUNDED to tf32 precision: adding
// half a tf32 ulp to the bit pattern turns the tensor core's truncation of lo into round-to-nearest. Truncating lo
// instead loses up to 2^-20 |v| always towards zero — a bias that adds up linearly over K and through the layers
// (measured: 10x the error of the fp32 SIMT path; with the rounding, ~2x).
__device__ __forceinline__ float tf32_lo(float v) {
    const float l = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    return __uint_as_float(__float_as_uint(l) + 0x1000u);
}


// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a wrong descriptor must abort the kernel (cudaErrorLaunchFailure), never hang the GPU.
#ifdef UF_TC_TIMING
__device__ long long g_tc_timing[16];  // per role: [0] total cycles, [1] cycles spent waiting (block 0, one thread per role)
#define TC_T0() long long t_role0 = clock64(), t_wait = 0
#define TC_WAIT(stmt) do { long long t_w = clock64(); stmt; t_wait += clock64() - t_w; } while (0)
#define TC_DONE(role) do { if (blockIdx.x == 0) { g_tc_timing[(role) * 2] = clock64() - t_role0; g_tc_timing[(role) * 2 + 1] = t_wait; } } while (0)
#else
#define TC_T0() do {} while (0)
#define TC_WAIT(stmt) stmt
#define TC_DONE(role) do {} while (0)
#endif

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    while (clock64() - t0 < 4000000000LL) {  // ~2 s at 1.9 GHz
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
    }
    printf("ultraface_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 32 (or 16) accumulator columns of this thread's TMEM lane, one wait for both loads
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v, bool second) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    if (second) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + 16));
    } else {
#pragma unroll
        for (int i = 16; i < 32; ++i) r[i] = 0u;
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// K-major, 128B-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows at a
// 128-byte pitch, 8-row groups 1024 bytes apart (SBO), version 1 (Blackwell), layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Same for a swizzle span of row_bytes = 64 or 128 (= the K extent of the tile): 8-row groups 8 * row_bytes apart,
// layout 4 (SWIZZLE_64B) / 2 (SWIZZLE_128B) of cute::UMMA::LayoutType.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : 4ull;
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)((8u * row_bytes) >> 4) << 32) | (1ull << 46) |
           (layout << 61);
}

constexpr int TC_BM = 128;        // rows per tile = UMMA M
constexpr int TC_BK = 32;         // fp32 per K block = one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB

struct TcParams {
    TView out, res;
    int has_res, relu;
    int M, K, N, n_umma, stages, tmem_cols;
    int tma_store;      // epilogue leaves through TMA (dense, 16-byte aligned rows, no residual)
    int wide;           // a_hi x [w_hi | w_lo] as ONE MMA into 2 * n_umma columns (the tensor pipe's time goes with the A
                        // operand read, M x K, not with N): 8 MMAs per K block instead of 12, the epilogue adds the halves
    // depthwise mode: the A operand is not loaded but COMPUTED by the converter warps, A = relu(dw3x3(in) + b)
    int dw_mode, dw_stride, dw_relu;
    TView dw_in;
    const float* dw_w;  // [9][K]
    const float* dw_b;  // [K]
    float bias[256];    // constant-bank operands of the epilogue
};

constexpr int TC_THREADS = 448;  // warps 0-3 converters, 4-7 converters (depthwise mode) or extra epilogue warps, 8-11 epilogue,
                                 // 12 TMA producer, 13 MMA issuer

__global__ void __launch_bounds__(TC_THREADS, 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_whi,
             const __grid_constant__ CUtensorMap tm_wlo, const __grid_constant__ CUtensorMap tm_out,
             const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_launch_dependents();
#ifdef UF_TC_TIMING
    const long long t_entry = clock64();
#endif
    // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the swizzle needs
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = p.n_umma * TC_BK * 4;
    const int stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                   // [stages] TMA landed
    uint64_t* conv = full + p.stages;        // [stages] A split into hi/lo
    uint64_t* empty = conv + p.stages;       // [stages] MMAs that read the stage are done
    uint64_t* acc_full = empty + p.stages;   // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint64_t* res_bar = acc_empty + 2;       // [8] residual tile landed (one per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (p.M + TC_BM - 1) / TC_BM;
    const int kblocks = (p.K + TC_BK - 1) / TC_BK;
    const int acc_stride = p.wide ? 2 * p.n_umma : p.n_umma;  // TMEM columns per accumulator buffer

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], p.dw_mode ? 8 : 4);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], (!p.dw_mode && p.tma_store) ? 8 : 4);
        }
        for (int w = 0; w < 8; ++w) mbar_init(&res_bar[w], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
#ifdef UF_TC_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 0) g_tc_timing[8] = clock64() - t_entry;
#endif
    pdl_wait();  // barriers, TMEM and the tensor-map fetch above overlap the predecessor; its results are needed from here

    if (warp == 12) {
        // ===== TMA producer =====
        if (lane == 0) {
            TC_T0();
            int s = 0;
            uint32_t ph = 0;
            const uint32_t tx_bytes = (uint32_t)((p.dw_mode ? 0 : TC_A_BYTES) + 2 * b_bytes);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    TC_WAIT(mbar_wait(&empty[s], ph ^ 1));
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full[s], tx_bytes);
                    if (!p.dw_mode) tma_load_2d(st, &tm_a, &full[s], kb * TC_BK, tile * TC_BM);
                    tma_load_2d(st + 2 * TC_A_BYTES, &tm_whi, &full[s], kb * TC_BK, 0);
                    tma_load_2d(st + 2 * TC_A_BYTES + b_bytes, &tm_wlo, &full[s], kb * TC_BK, 0);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
            TC_DONE(0);
        }
    } else if (warp == 13) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major both, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_umma >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * p.n_umma) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            TC_T0();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int a = it & 1;
                const uint32_t aph = (uint32_t)((it >> 1) & 1);
                TC_WAIT(mbar_wait(&acc_empty[a], aph ^ 1));
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(a * acc_stride);
                uint32_t accumulate = 0;
                for (int kb = 0; kb < kblocks; ++kb) {
                    TC_WAIT(mbar_wait(&full[s], ph));
                    TC_WAIT(mbar_wait(&conv[s], ph));
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                    // descriptors once per stage; advancing K by 8 tf32 = 32 bytes is +2 in the 16-byte address field
                    const uint64_t d_ahi = umma_desc_sw128(st), d_alo = umma_desc_sw128(st + TC_A_BYTES);
                    const uint64_t d_whi = umma_desc_sw128(st + 2 * TC_A_BYTES), d_wlo = umma_desc_sw128(st + 2 * TC_A_BYTES + b_bytes);
                    if (p.wide) {  // W_hi and W_lo are adjacent in the stage: one B operand of 2 * n_umma rows
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) { umma_tf32(d, d_ahi + 2 * k, d_whi + 2 * k, idesc2, accumulate); accumulate = 1; }
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(d, d_alo + 2 * k, d_whi + 2 * k, idesc, 1);
                    } else {
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) { umma_tf32(d, d_alo + 2 * k, d_whi + 2 * k, idesc, accumulate); accumulate = 1; }
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(d, d_ahi + 2 * k, d_wlo + 2 * k, idesc, 1);
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) umma_tf32(d, d_ahi + 2 * k, d_whi + 2 * k, idesc, 1);
                    }
                    umma_commit(&empty[s]);  // implies tcgen05.fence::before_thread_sync
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                umma_commit(&acc_full[a]);
            }
            TC_DONE(1);
        }
    } else if (warp < 8 && p.dw_mode) {
        // ===== converters, depthwise mode: A[row][k] = relu(dw3x3(in)[pixel row][channel k] + b) is computed here
        // (the depthwise result never exists in memory) and written as (raw = tf32 hi, lo) in the 128B-swizzled
        // K-major layout the UMMA descriptors expect: 16-byte chunk q of row r lives at r*128 + ((q ^ (r & 7)) << 4).
        TC_T0();
        const int ctid = threadIdx.x;          // 0..255
        const int q = ctid & 7;                // 4-channel chunk of the 32-channel K block
        const int r0 = ctid >> 3;              // rows r0 + 32*j, j = 0..3
        const TView in = p.dw_in;
        const int S = p.dw_stride, Wo = p.out.W, HWo = p.out.H * p.out.W;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            long long base[4];
            int iy0[4], ix0[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = tile * TC_BM + r0 + 32 * j;
                if (m < p.M) {
                    const int f = m / HWo, pix = m - f * HWo;
                    const int y = pix / Wo, x = pix - y * Wo;
                    base[j] = (long long)f * in.frame_stride;
                    iy0[j] = y * S - 1; ix0[j] = x * S - 1;
                } else {
                    base[j] = -1; iy0[j] = 0; ix0[j] = 0;
                }
            }
            for (int kb = 0; kb < kblocks; ++kb) {
                TC_WAIT(mbar_wait(&empty[s], ph ^ 1));  // the MMAs that read this stage have finished
                const int c = kb * TC_BK + q * 4;
                float4 wv[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) wv[t] = __ldg(reinterpret_cast<const float4*>(p.dw_w + (size_t)t * p.K + c));
                const float4 bias4 = __ldg(reinterpret_cast<const float4*>(p.dw_b + c));
                uint8_t* a_hi = smem + (size_t)s * stage_bytes;
                uint8_t* a_lo = a_hi + TC_A_BYTES;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {  // not unrolled: keeps 9 weights + one item's taps live, nothing more
                    const int r = r0 + 32 * j;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (base[j] >= 0) {
                        acc = bias4;
                        const float* ip = in.p + base[j] + c;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {  // one input row at a time: 3 loads in flight, few live registers
                            const int iy = iy0[j] + ky;
                            const bool rok = iy >= 0 && iy < in.H;
                            float4 v[3];
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const int ix = ix0[j] + kx;
                                v[kx] = (rok && ix >= 0 && ix < in.W)
                                            ? *reinterpret_cast<const float4*>(ip + ((size_t)iy * in.W + ix) * in.pix_stride)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float4 ww = wv[ky * 3 + kx];
                                acc.x = fmaf(v[kx].x, ww.x, acc.x); acc.y = fmaf(v[kx].y, ww.y, acc.y);
                                acc.z = fmaf(v[kx].z, ww.z, acc.z); acc.w = fmaf(v[kx].w, ww.w, acc.w);
                            }
                        }
                        if (p.dw_relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
                    }
                    float4 l;
                    l.x = tf32_lo(acc.x);
                    l.y = tf32_lo(acc.y);
                    l.z = tf32_lo(acc.z);
                    l.w = tf32_lo(acc.w);
                    const int off = r * 128 + ((q ^ (r & 7)) << 4);
                    *reinterpret_cast<float4*>(a_hi + off) = acc;  // kind::tf32 ignores the low 13 mantissa bits: raw = hi
                    *reinterpret_cast<float4*>(a_lo + off) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        if (threadIdx.x == 0) TC_DONE(2);
    } else if (warp < 4 && !p.dw_mode) {
        // ===== converters: fp32 -> (tf32 hi, tf32 lo), element-wise so the swizzled layout is preserved =====
        TC_T0();
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < kblocks; ++kb) {
                TC_WAIT(mbar_wait(&full[s], ph));
                float4* hi = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + TC_A_BYTES);
#pragma unroll
                for (int j = 0; j < TC_A_BYTES / 16 / 128; ++j) {
                    const int i = threadIdx.x + j * 128;
                    const float4 v = hi[i];
                    float4 l;
                    l.x = tf32_lo(v.x);
                    l.y = tf32_lo(v.y);
                    l.z = tf32_lo(v.z);
                    l.w = tf32_lo(v.w);
                    // hi is not written back: kind::tf32 ignores the low 13 mantissa bits of its operands, i.e. the
                    // raw fp32 tile already *is* a_hi (checked by the 1e-4 layer-parity tests)
                    lo[i] = l;
                }
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        if (threadIdx.x == 0) TC_DONE(2);
    } else if (warp < 8 && !(!p.dw_mode && p.tma_store)) {
        // idle: these warps are converters in depthwise mode and extra epilogue warps with the TMA-store epilogue
    } else {
        // ===== epilogue warps 8..11: TMEM -> registers -> per-warp smem transpose -> coalesced global =====
        // A TMEM lane is an output row, so a lane owns a whole row; writing rows straight from the
        // lanes would touch 32 different cache lines per store. Each warp stages 32 rows x 32 columns
        // in its own padded smem tile and writes them back 4 rows per instruction (8 lanes x 16 B =
        // one 128-byte line per row), with bias / residual / ReLU applied in the coalesced phase.
        const int q = warp & 3;
        if (p.tma_store) {
            // fast path: +bias (constant bank) / residual / ReLU in registers -> 128B-swizzled 32x32 tile per warp ->
            // TMA store (rows past M and columns past N are clipped by the tensor map): no address arithmetic at all.
            // In plain mode warps 4-7 join as a second epilogue group: a TMEM lane quarter (32 rows) is then shared by
            // two warps that take alternate 32-column chunks.
            const int ngrp = p.dw_mode ? 1 : 2;
            const int grp = p.dw_mode ? 0 : (warp >= 8 ? 1 : 0);
            const int ew = grp * 4 + q;  // epilogue warp index 0..7
            uint8_t* st = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 1023) & ~uintptr_t(1023)) + ew * 4096;
            uint32_t rph = 0;
            TC_T0();
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int a = it & 1;
                const uint32_t aph = (uint32_t)((it >> 1) & 1);
                TC_WAIT(mbar_wait(&acc_full[a], aph));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * acc_stride);
                bool released = false;
                for (int c0 = 32 * grp; c0 < p.n_umma; c0 += 32 * ngrp) {
                    if (lane == 0) bulk_wait_read0();  // this warp's previous store has left its staging tile
                    __syncwarp();
                    if (p.has_res && lane == 0) {  // residual tile by TMA into the staging tile (rows past M read as zero)
                        mbar_expect_tx(&res_bar[ew], 4096);
                        tma_load_2d(st, &tm_res, &res_bar[ew], c0, tile * TC_BM + q * 32);
                    }
                    float v[32];
                    tmem_ld32(taddr + c0, v, c0 + 16 < p.n_umma);
                    if (p.wide) {  // + the a_hi * w_lo half
                        float v2[32];
                        tmem_ld32(taddr + p.n_umma + c0, v2, c0 + 16 < p.n_umma);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += v2[j];
                    }
                    if (c0 + 32 * ngrp >= p.n_umma) {  // last read of this accumulator: hand it back to the MMA warp now
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[a]);
                        released = true;
                    }
                    if (p.has_res) {
                        mbar_wait(&res_bar[ew], rph);
                        rph ^= 1;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 r4 = *reinterpret_cast<const float4*>(st + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4));
                            v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
                        }
                        __syncwarp();  // every lane has read its residual row before the tile is overwritten
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 x;
                        x.x = v[j] + p.bias[(c0 + j) & 255]; x.y = v[j + 1] + p.bias[(c0 + j + 1) & 255];
                        x.z = v[j + 2] + p.bias[(c0 + j + 2) & 255]; x.w = v[j + 3] + p.bias[(c0 + j + 3) & 255];
                        if (p.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                        *reinterpret_cast<float4*>(st + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) = x;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tm_out, st, c0, tile * TC_BM + q * 32);
                        bulk_commit();
                    }
                }
                if (!released) {  // this warp had no column chunk in this tile (N <= 32): still release the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[a]);
                }
            }
            if (lane == 0) bulk_wait0();
            if (threadIdx.x == 256) TC_DONE(3);
        } else {
                float* stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15)) + q * (32 * 36);
        const int HW = p.out.H * p.out.W;
        const bool vec = ((p.N & 3) == 0) && ((p.out.pix_stride & 3) == 0) && ((reinterpret_cast<size_t>(p.out.p) & 15) == 0) &&
                         ((p.out.frame_stride & 3) == 0) &&
                         (!p.has_res || (((p.res.pix_stride & 3) == 0) && ((reinterpret_cast<size_t>(p.res.p) & 15) == 0) &&
                                         ((p.res.frame_stride & 3) == 0)));
        const int cl = (lane & 7) * 4;   // column (within the 32-column chunk) this lane handles when storing
        const int rl = lane >> 3;        // row (within a group of 4) this lane handles when storing
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const uint32_t aph = (uint32_t)((it >> 1) & 1);
            mbar_wait(&acc_full[a], aph);
            tc_fence_after();
            const int m_base = tile * TC_BM + q * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * acc_stride);
            for (int c0 = 0; c0 < p.n_umma; c0 += 32) {
                float v[32];
                tmem_ld16(taddr + c0, v);  // warp-collective: executed by every lane
                if (c0 + 16 < p.n_umma) tmem_ld16(taddr + c0 + 16, v + 16);
                if (p.wide) {
                    float v2[32];
                    tmem_ld32(taddr + p.n_umma + c0, v2, c0 + 16 < p.n_umma);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += (j < 16 || c0 + 16 < p.n_umma) ? v2[j] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                const int n = c0 + cl;
                float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < p.N) {
                    bq.x = p.bias[n & 255];
                    if (n + 1 < p.N) bq.y = p.bias[(n + 1) & 255];
                    if (n + 2 < p.N) bq.z = p.bias[(n + 2) & 255];
                    if (n + 3 < p.N) bq.w = p.bias[(n + 3) & 255];
                }
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += 4) {
                    const int r = r0 + rl;
                    const int m = m_base + r;
                    if (m < p.M && n < p.N) {
                        const int f = m / HW;
                        const int pix = m - f * HW;
                        float4 x = *reinterpret_cast<const float4*>(stage + r * 36 + cl);
                        x.x += bq.x; x.y += bq.y; x.z += bq.z; x.w += bq.w;
                        float* op = p.out.p + (size_t)f * p.out.frame_stride + (size_t)pix * p.out.pix_stride + n;
                        const float* rp = p.has_res ? p.res.p + (size_t)f * p.res.frame_stride + (size_t)pix * p.res.pix_stride + n : nullptr;
                        if (vec && n + 4 <= p.N) {
                            if (rp) { const float4 rr = *reinterpret_cast<const float4*>(rp); x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w; }
                            if (p.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                            *reinterpret_cast<float4*>(op) = x;
                        } else {
                            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (n + j < p.N) {
                                    float y = xv[j] + (rp ? rp[j] : 0.f);
                                    op[j] = p.relu ? fmaxf(y, 0.f) : y;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef UF_TC_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 0) g_tc_timing[9] = clock64() - t_entry;
#endif
    if (warp == 13) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so the
// library does not link libcuda) and the launcher.
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp32 matrix [rows][cols] with a row pitch of row_stride_bytes; box = box_rows x 32 floats, 128B swizzle,
// out-of-bounds elements read as zero.
bool make_tmap_f32_2d(TmaMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows) {
    static_assert(sizeof(TmaMap) == sizeof(CUtensorMap), "TmaMap must mirror CUtensorMap");
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {TC_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim,
                    gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// K4+K5 fused, TMA-pipelined persistent form for the large memory-bound maps (C = 16/32/64):
//   * the input tile (+halo) arrives by TMA as 16-channel slices [IH][IW][16] (64B swizzle,
//     out-of-bounds = zero padding for free), double-buffered behind mbarriers, so the loads of
//     the next slice / next tile overlap the arithmetic of the current one;
//   * thread = output pixel: depthwise 3x3 per slice into registers (C values), then the 1x1 conv
//     against weights broadcast from shared memory, 32 outputs per pass;
//   * results are staged in a 128B-swizzled tile and leave by TMA store (edge tiles clipped by the
//     tensor map), so neither loads nor stores cost address arithmetic or uncoalesced sectors.
// ---------------------------------------------------------------------------------------------
// Weights travel as a __grid_constant__ kernel parameter: with the loops fully unrolled every FFMA takes
// its weight as a constant-bank operand, so the 1x1 conv costs no shared-memory bandwidth at all
// (broadcasting weights from smem costs one wavefront per weight per warp and caps FFMA at 25 %).
template <int C, int N>
struct FusedWeights {
    float dw[9 * C];   // [tap][c]
    float dwb[C];
    float pw[C * N];   // [ci][n]
    float pwb[N];
};

template <int C, int N, int S, int TY>
__global__ void __launch_bounds__(8 * TY, (8 * TY == 256) ? 2 : 3)
fused_dwpw_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                      const __grid_constant__ FusedWeights<C, N> wts, int dw_relu, int pw_relu, int tiles_x, int tiles_y,
                      int total_tiles) {
    constexpr int TX = 8, IW = (TX - 1) * S + 3, IH = (TY - 1) * S + 3, NSL = C / 16, NTHR = TX * TY;
    constexpr int SLICE_BYTES = IH * IW * 64;
    constexpr int SLICE_ALLOC = (SLICE_BYTES + 1023) / 1024 * 1024;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* in_buf = smem;                                   // 2 x SLICE_ALLOC
    uint8_t* out_stage = smem + 2 * SLICE_ALLOC;              // NTHR x 128 B
    uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + NTHR * 128);  // [2]
    const int tid = threadIdx.x;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_units = my_tiles * NSL;

    auto tile_coords = [&](int t, int& f, int& y0, int& x0) {
        int b = blockIdx.x + t * gridDim.x;
        const int txi = b % tiles_x; b /= tiles_x;
        const int tyi = b % tiles_y;
        f = b / tiles_y;
        x0 = txi * TX; y0 = tyi * TY;
    };
    auto issue = [&](int u) {  // thread 0 only
        int f, y0, x0;
        tile_coords(u / NSL, f, y0, x0);
        uint64_t* bar = &full[u & 1];
        mbar_expect_tx(bar, SLICE_BYTES);
        tma_load_4d(in_buf + (u & 1) * SLICE_ALLOC, &tm_in, bar, (u % NSL) * 16, x0 * S - 1, y0 * S - 1, f);
    };

    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();
    __syncthreads();
    if (tid == 0 && n_units > 0) {
        issue(0);
        if (n_units > 1) issue(1);
    }
    const int tx = tid % TX, ty = tid / TX;
    int u = 0;
    for (int t = 0; t < my_tiles; ++t) {
        float d[C];
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl, ++u) {
            mbar_wait(&full[u & 1], (uint32_t)((u >> 1) & 1));
            const uint8_t* buf = in_buf + (u & 1) * SLICE_ALLOC;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = sl * 16 + q * 4;
                float a0 = wts.dwb[c], a1 = wts.dwb[c + 1], a2 = wts.dwb[c + 2], a3 = wts.dwb[c + 3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int r = (ty * S + ky) * IW + tx * S + kx;  // 64-byte row of the box
                        const float4 v = *reinterpret_cast<const float4*>(buf + r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
                        const int wo = (ky * 3 + kx) * C + c;
                        a0 = fmaf(v.x, wts.dw[wo], a0); a1 = fmaf(v.y, wts.dw[wo + 1], a1);
                        a2 = fmaf(v.z, wts.dw[wo + 2], a2); a3 = fmaf(v.w, wts.dw[wo + 3], a3);
                    }
                if (dw_relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
                d[c] = a0; d[c + 1] = a1; d[c + 2] = a2; d[c + 3] = a3;
            }
            if (tid == 0 && sl == NSL - 1) bulk_wait_read0();  // the previous tile's store has left the staging tile
            __syncthreads();                                    // everyone is done with this slice buffer
            if (tid == 0 && u + 2 < n_units) issue(u + 2);
        }
        int f, y0, x0;
        tile_coords(t, f, y0, x0);
#pragma unroll
        for (int n0 = 0; n0 < N; n0 += 32) {
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = wts.pwb[n0 + j];
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = fmaf(d[ci], wts.pw[ci * N + n0 + j], o[j]);
            }
            if (n0 > 0) {  // second pass of a 64-output layer: wait until the first pass' store has read the tile
                if (tid == 0) bulk_wait_read0();
                __syncthreads();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 v = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                if (pw_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4*>(out_stage + tid * 128 + ((q ^ (tid & 7)) << 4)) = v;  // 128B swizzle
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tma_store_4d(&tm_out, out_stage, n0, x0, y0, f);
                bulk_commit();
            }
        }
    }
    if (tid == 0) bulk_wait0();  // smem must outlive the last store
}

// ---------------------------------------------------------------------------------------------
// Tensor-core form of the kernel above: same TMA-fed depthwise front half (thread = output pixel, 16-channel slices,
// depthwise weights as constant operands), but the 1x1 conv — 78 % of the kernel's FFMAs — leaves the SIMT pipes:
// every thread writes its pixel's C depthwise results as one K-major row of the A operand (raw fp32 = tf32 hi, and
// the lo remainder) in the canonical swizzled layout, one thread issues the 3xTF32 tcgen05.mma sequence
// (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi) against the CTA-resident W_hi / W_lo tiles into a TMEM accumulator, and the
// same threads read their row back (tcgen05.ld: warp w owns TMEM lanes 32w..32w+31 = its own pixels), add the bias,
// apply ReLU and stage it for the TMA store. CTA = 8 x 16 pixels = one UMMA M tile of 128 rows; the output staging
// tile aliases the A tiles (dead once the MMAs have committed), ~70 KB of shared memory, three CTAs per SM cover
// each other's MMA / TMA latencies.
// ---------------------------------------------------------------------------------------------
template <int C>
struct DwWeights {
    float dw[9 * C];  // [tap][c]
    float dwb[C];
    float pwb[64];    // bias of the 1x1 conv (N <= 64)
};

// A kernel that allocates tensor memory runs ONE CTA per SM (cudaOccupancyMaxActiveBlocksPerMultiprocessor = 1
// whatever its footprint: the CTA owns the SM's TMEM), and a single 128-thread tile pipeline leaves the SM idle during
// every barrier, MMA round trip and TMA wait (measured: 1.9-2.7 us per tile, slower than the SIMT 1x1). So the CTA holds
// G independent GROUPS of 128 threads, each a complete tile pipeline of its own — input slices, A / staging buffer,
// mbarriers, accumulator columns, a named barrier (bar.sync 1+g, 128) — sharing only the resident W tiles: what
// three or four co-resident CTAs would have been.
template <int C, int N, int S, int G, int NB>
__global__ void __launch_bounds__(128 * G, 1)
fused_dwpw_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                     const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo,
                     const __grid_constant__ DwWeights<C> wts, int dw_relu, int pw_relu, int tiles_x, int tiles_y,
                     int total_tiles) {
    constexpr int TX = 8, TY = 16, IW = (TX - 1) * S + 3, IH = (TY - 1) * S + 3;
    constexpr int SC = 16;                   // channels per input slice (64-byte box rows; 32-channel / 128-byte slices were
    constexpr int NSL = C / SC;              // measured: no faster per tile, and they cost a group of shared memory)
    constexpr int SROW = SC * 4, SQ = SC / 4;
    constexpr int SLICE_BYTES = IH * IW * SROW;
    constexpr int SLICE_ALLOC = (SLICE_BYTES + 1023) / 1024 * 1024;
    constexpr int KBLK = C > 32 ? C / 32 : 1;  // K blocks of the operand tiles: one swizzle span (<= 32 channels) each
    constexpr int KC = C / KBLK;             // channels per K block (16 or 32)
    constexpr int ROWB = KC * 4;             // bytes per A / W row within a K block = swizzle span (64 or 128)
    constexpr int AKB = 128 * ROWB;          // one K block of an A tile (8 or 16 KB)
    constexpr int WKB = N * ROWB;            // one K block of a W tile
    constexpr int A_BYTES = KBLK * AKB;
    constexpr int A_REGION = 2 * A_BYTES < 16384 ? 16384 : 2 * A_BYTES;  // A_hi | A_lo, and at least the 16 KB staging tile
    constexpr int GROUP_BYTES = NB * SLICE_ALLOC + A_REGION;  // NB input slices in flight per group (memory latency)
    constexpr int W_BYTES = KBLK * WKB;
    constexpr int NPAD = N <= 32 ? 32 : 64;  // N as the MMA sees it
    // One K block: W_hi and W_lo sit back to back in shared memory = ONE B operand of 2N rows, so a_hi * [w_hi | w_lo]
    // is a single MMA into 2N accumulator columns (these small MMAs cost the same whatever their N) and a_lo * w_hi a
    // second one into the first N: two MMAs per k-step instead of three, the epilogue adds the two halves.
    constexpr bool WIDE = KBLK == 1 && N == NPAD;
    constexpr int NCOL = WIDE ? 2 * NPAD : NPAD;  // accumulator columns per group
    constexpr uint32_t TMEM_COLS = G * NCOL <= 32 ? 32 : (G * NCOL <= 64 ? 64 : (G * NCOL <= 128 ? 128 : (G * NCOL <= 256 ? 256 : 512)));
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, g = tid >> 7, lt = tid & 127, lwarp = lt >> 5;
    uint8_t* in_buf = smem + g * GROUP_BYTES;                 // NB x SLICE_ALLOC
    uint8_t* a_hi = in_buf + NB * SLICE_ALLOC;                 // A_hi | A_lo; the staging tile (128 x 128 B) aliases them
    uint8_t* a_lo = a_hi + A_BYTES;
    uint8_t* out_stage = a_hi;
    uint8_t* w_hi = smem + G * GROUP_BYTES;
    uint8_t* w_lo = w_hi + W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_lo + W_BYTES);
    uint64_t* w_bar = bars;
    uint64_t* full = bars + 1 + g * (NB + 1);                 // [NB] input slices of this group
    uint64_t* mma_bar = full + NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + G * (NB + 1));
    const int vcta = blockIdx.x * G + g, vgrid = gridDim.x * G;  // this group's place among all tile pipelines
    const int my_tiles = vcta < total_tiles ? (total_tiles - vcta + vgrid - 1) / vgrid : 0;
    const int n_units = my_tiles * NSL;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };

    auto tile_coords = [&](int t, int& f, int& y0, int& x0) {
        int b = vcta + t * vgrid;
        const int txi = b % tiles_x; b /= tiles_x;
        const int tyi = b % tiles_y;
        f = b / tiles_y;
        x0 = txi * TX; y0 = tyi * TY;
    };
    auto issue = [&](int u) {  // group leader only
        int f, y0, x0;
        tile_coords(u / NSL, f, y0, x0);
        uint64_t* bar = &full[u % NB];
        mbar_expect_tx(bar, SLICE_BYTES);
        tma_load_4d(in_buf + (u % NB) * SLICE_ALLOC, &tm_in, bar, (u % NSL) * SC, x0 * S - 1, y0 * S - 1, f);
    };

    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(w_bar, 1);
        for (int i = 0; i < G * (NB + 1); ++i) mbar_init(bars + 1 + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {  // the weights are static: fetched before the predecessor kernel is waited for
        mbar_expect_tx(w_bar, 2 * W_BYTES);
        for (int kb = 0; kb < KBLK; ++kb) {
            tma_load_2d(w_hi + kb * WKB, &tm_whi, w_bar, kb * KC, 0);
            tma_load_2d(w_lo + kb * WKB, &tm_wlo, w_bar, kb * KC, 0);
        }
    }
    pdl_wait();
    if (lt == 0) {
        for (int i = 0; i < NB && i < n_units; ++i) issue(i);
    }
    const int tx = lt % TX, ty = lt / TX;
    // instruction descriptor: D=F32, A=B=TF32, K-major both, N>>3, M>>4 (M = 128)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * NPAD) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // this thread's row of the A tiles: 16-byte chunk q lives at q ^ (row & 7) (128B swizzle) or q ^ ((row >> 1) & 3) (64B)
    const int a_row = lt * ROWB;
    const int a_swz = ROWB == 128 ? (lt & 7) : ((lt >> 1) & 3);
    const uint32_t acc = tmem_base + (uint32_t)(g * NCOL);                 // this group's accumulator columns
    const uint32_t taddr = acc + ((uint32_t)(lwarp * 32) << 16);          // warp w of a group owns TMEM lanes 32w..32w+31
    int u = 0;
    for (int t = 0; t < my_tiles; ++t) {
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl, ++u) {
            mbar_wait(&full[u % NB], (uint32_t)((u / NB) & 1));
            const uint8_t* buf = in_buf + (u % NB) * SLICE_ALLOC;
            float4 dq[SQ];
#pragma unroll
            for (int q = 0; q < SQ; ++q) {
                const int c = sl * SC + q * 4;
                float a0 = wts.dwb[c], a1 = wts.dwb[c + 1], a2 = wts.dwb[c + 2], a3 = wts.dwb[c + 3];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int r = (ty * S + ky) * IW + tx * S + kx;  // row of the box (one pixel's SC channels)
                        const int sw = SROW == 128 ? (r & 7) : ((r >> 1) & 3);
                        const float4 v = *reinterpret_cast<const float4*>(buf + r * SROW + ((q ^ sw) << 4));
                        const int wo = (ky * 3 + kx) * C + c;
                        a0 = fmaf(v.x, wts.dw[wo], a0); a1 = fmaf(v.y, wts.dw[wo + 1], a1);
                        a2 = fmaf(v.z, wts.dw[wo + 2], a2); a3 = fmaf(v.w, wts.dw[wo + 3], a3);
                    }
                if (dw_relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
                dq[q] = make_float4(a0, a1, a2, a3);
            }
            if (sl == 0) {  // first write into the A buffer: the previous tile's store (staged there) must have read it
                if (lt == 0) bulk_wait_read0();
                group_sync();
            }
#pragma unroll
            for (int q = 0; q < SQ; ++q) {
                const int ch = sl * SQ + q;                 // 16-byte chunk of the pixel's C channels
                const int off = (ch / (KC / 4)) * AKB + a_row + (((ch % (KC / 4)) ^ a_swz) << 4);
                float4 l;
                l.x = tf32_lo(dq[q].x);
                l.y = tf32_lo(dq[q].y);
                l.z = tf32_lo(dq[q].z);
                l.w = tf32_lo(dq[q].w);
                *reinterpret_cast<float4*>(a_hi + off) = dq[q];  // kind::tf32 ignores the low 13 mantissa bits: raw = hi
                *reinterpret_cast<float4*>(a_lo + off) = l;
            }
            if (sl == NSL - 1) fence_proxy_async();  // A rows (generic proxy) -> visible to the tensor core
            group_sync();                            // the group is done with this slice buffer (and, last slice, with A)
            if (lt == 0 && u + NB < n_units) issue(u + NB);
        }
        if (lt == 0) {
            if (t == 0) mbar_wait(w_bar, 0);
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < KBLK; ++kb) {
                const uint64_t d_ahi = umma_desc_kmajor(smem_u32(a_hi + kb * AKB), ROWB), d_alo = umma_desc_kmajor(smem_u32(a_lo + kb * AKB), ROWB);
                const uint64_t d_whi = umma_desc_kmajor(smem_u32(w_hi + kb * WKB), ROWB), d_wlo = umma_desc_kmajor(smem_u32(w_lo + kb * WKB), ROWB);
                if (WIDE) {
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k) umma_tf32(acc, d_ahi + 2 * k, d_whi + 2 * k, idesc2, k > 0 ? 1u : 0u);  // [w_hi | w_lo]
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k) umma_tf32(acc, d_alo + 2 * k, d_whi + 2 * k, idesc, 1);
                } else {
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k) umma_tf32(acc, d_alo + 2 * k, d_whi + 2 * k, idesc, (kb | k) > 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k) umma_tf32(acc, d_ahi + 2 * k, d_wlo + 2 * k, idesc, 1);
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k) umma_tf32(acc, d_ahi + 2 * k, d_whi + 2 * k, idesc, 1);
                }
            }
            umma_commit(mma_bar);
        }
        __syncwarp();
        mbar_wait(mma_bar, (uint32_t)(t & 1));
        tc_fence_after();
        int f, y0, x0;
        tile_coords(t, f, y0, x0);
#pragma unroll
        for (int n0 = 0; n0 < N; n0 += 32) {
            float v[32];
            tmem_ld32(taddr + n0, v, true);
            if (WIDE) {  // + the a_hi * w_lo half
                float v2[32];
                tmem_ld32(taddr + NPAD + n0, v2, true);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += v2[j];
            }
            if (n0 > 0) {  // second pass of a 64-output layer: wait until the first pass' store has read the staging tile
                if (lt == 0) bulk_wait_read0();
                group_sync();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 x = make_float4(v[q * 4] + wts.pwb[n0 + q * 4], v[q * 4 + 1] + wts.pwb[n0 + q * 4 + 1],
                                       v[q * 4 + 2] + wts.pwb[n0 + q * 4 + 2], v[q * 4 + 3] + wts.pwb[n0 + q * 4 + 3]);
                if (pw_relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                *reinterpret_cast<float4*>(out_stage + lt * 128 + ((q ^ (lt & 7)) << 4)) = x;  // 128B swizzle
            }
            fence_proxy_async();
            tc_fence_before();  // this thread's tcgen05.ld are complete: the next tile's MMAs may overwrite the accumulator
            group_sync();
            if (lt == 0) {
                tma_store_4d(&tm_out, out_stage, n0, x0, y0, f);
                bulk_commit();
            }
        }
    }
    if (lt == 0) bulk_wait0();  // smem must outlive the last store
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// NHWC fp32 view as a 4-D tensor map (C, W, H, frames); box = (box_c, box_w, box_h, 1)
bool make_tmap_nhwc(TmaMap* out, const TView& v, int frames, uint32_t box_c, uint32_t box_w, uint32_t box_h, int swizzle_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)frames};
    cuuint64_t gstride[3] = {(cuuint64_t)v.pix_stride * 4, (cuuint64_t)v.W * v.pix_stride * 4, (cuuint64_t)v.frame_stride * 4};
    cuuint32_t box[4] = {box_c, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, v.p, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool fused_dwpw_tma_supported(int C, int N, int stride) {
    if (C == 16) return N == 32 && stride == 1;
    return C == 32 && (N == 32 || N == 64) && (stride == 1 || stride == 2);
}

// the tcgen05 form also takes the 64-channel pairs of the 40x30 map (two K blocks per operand tile)
bool fused_dwpw_tc_supported(int C, int N, int stride) {
    return fused_dwpw_tma_supported(C, N, stride) || (C == 64 && N == 64 && stride == 1);
}

void fused_dwpw_tma_boxes(int stride, int* in_w, int* in_h, int* out_w, int* out_h) {
    const int TX = 8, TY = stride == 1 ? 32 : 16;
    *in_w = (TX - 1) * stride + 3; *in_h = (TY - 1) * stride + 3; *out_w = TX; *out_h = TY;
}

template <int C, int N, int S, int TY>
static void launch_tma_t(const TmaMap& tm_in, const TmaMap& tm_out, const TView& out, const float* host_w, int dw_relu,
                         int pw_relu, int frames, cudaStream_t s) {
    constexpr int TX = 8, IW = (TX - 1) * S + 3, IH = (TY - 1) * S + 3;
    constexpr int SLICE_ALLOC = (IH * IW * 64 + 1023) / 1024 * 1024;
    static_assert(sizeof(FusedWeights<C, N>) + 2 * sizeof(CUtensorMap) + 64 <= 32764, "kernel parameter space exceeded");
    const int tiles_x = (out.W + TX - 1) / TX, tiles_y = (out.H + TY - 1) / TY;
    const int total = tiles_x * tiles_y * frames;
    const size_t smem = 2 * SLICE_ALLOC + TX * TY * 128 + 16 + 1024;
    auto kern = fused_dwpw_tma_kernel<C, N, S, TY>;
    static bool configured[64] = {};
    static int ctas_per_sm[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int nb = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, TX * TY, smem);
        ctas_per_sm[dev & 63] = nb < 1 ? 1 : nb;
        configured[dev & 63] = true;
    }
    int grid = 148 * ctas_per_sm[dev & 63];
    if (grid > total) grid = total;
    launch_pdl(kern, dim3(grid), dim3(TX * TY), smem, s, *reinterpret_cast<const CUtensorMap*>(&tm_in),
               *reinterpret_cast<const CUtensorMap*>(&tm_out), *reinterpret_cast<const FusedWeights<C, N>*>(host_w), dw_relu, pw_relu,
               tiles_x, tiles_y, total);
}

size_t fused_dwpw_tma_weight_floats(int C, int N) { return (size_t)10 * C + (size_t)C * N + N; }

// host_w: [dw 9*C][dw bias C][pw C*N][pw bias N] on the HOST (it is copied into the kernel parameter space)
void launch_fused_dwpw_tma(const TmaMap& tm_in, const TmaMap& tm_out, const TView& in, const TView& out, const float* host_w,
                           int stride, int dw_relu, int pw_relu, int frames, cudaStream_t s) {
#define UF_T(CC, NN, SS, TYY) launch_tma_t<CC, NN, SS, TYY>(tm_in, tm_out, out, host_w, dw_relu, pw_relu, frames, s)
    const int C = in.C, N = out.C;
    if (C == 16 && N == 32 && stride == 1) UF_T(16, 32, 1, 32);
    else if (C == 32 && N == 32 && stride == 1) UF_T(32, 32, 1, 32);
    else if (C == 32 && N == 32 && stride == 2) UF_T(32, 32, 2, 16);
    else if (C == 32 && N == 64 && stride == 2) UF_T(32, 64, 2, 16);
    else if (C == 32 && N == 64 && stride == 1) UF_T(32, 64, 1, 32);
#undef UF_T
}

// fp32 matrix [rows][cols], box = box_rows x box_cols floats, swizzle span = box_cols * 4 bytes (64 or 128)
bool make_tmap_f32_2d_sw(TmaMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                         uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = box_cols * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim,
                    gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int fused_dwpw_tc_slice_channels(int) { return 16; }

void fused_dwpw_tc_boxes(int stride, int* in_w, int* in_h, int* out_w, int* out_h) {
    const int TX = 8, TY = 16;
    *in_w = (TX - 1) * stride + 3; *in_h = (TY - 1) * stride + 3; *out_w = TX; *out_h = TY;
}

template <int C, int N, int S, int G, int NB>
static void launch_fused_tc_t(const TmaMap& tm_in, const TmaMap& tm_out, const TmaMap& tm_whi, const TmaMap& tm_wlo,
                              const TView& out, const float* host_w, int dw_relu, int pw_relu, int frames, cudaStream_t s) {
    constexpr int TX = 8, TY = 16, IW = (TX - 1) * S + 3, IH = (TY - 1) * S + 3;
    constexpr int SLICE_ALLOC = (IH * IW * 64 + 1023) / 1024 * 1024;
    constexpr int A_REGION = 2 * 128 * C * 4 < 16384 ? 16384 : 2 * 128 * C * 4;
    constexpr size_t SMEM = (size_t)G * (NB * SLICE_ALLOC + A_REGION) + 2 * N * C * 4 + (2 + (NB + 1) * G) * 8 + 16 + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
    static_assert(sizeof(DwWeights<C>) + 4 * sizeof(CUtensorMap) + 64 <= 32764, "kernel parameter space exceeded");
    DwWeights<C> w;
    memcpy(w.dw, host_w, sizeof(float) * 10 * C);  // dw taps followed by the dw bias
    for (int i = 0; i < 64; ++i) w.pwb[i] = i < N ? host_w[10 * C + C * N + i] : 0.f;
    const int tiles_x = (out.W + TX - 1) / TX, tiles_y = (out.H + TY - 1) / TY;
    const int total = tiles_x * tiles_y * frames;
    auto kern = fused_dwpw_tc_kernel<C, N, S, G, NB>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured[dev & 63] = true;
    }
    int grid = 148;  // one CTA per SM (tensor memory), G tile pipelines each
    if (grid > (total + G - 1) / G) grid = (total + G - 1) / G;
    launch_pdl(kern, dim3(grid), dim3(128 * G), SMEM, s, *reinterpret_cast<const CUtensorMap*>(&tm_in),
               *reinterpret_cast<const CUtensorMap*>(&tm_out), *reinterpret_cast<const CUtensorMap*>(&tm_whi),
               *reinterpret_cast<const CUtensorMap*>(&tm_wlo), w, dw_relu, pw_relu, tiles_x, tiles_y, total);
}

// host_w as for launch_fused_dwpw_tma; tm_whi / tm_wlo: the 1x1 weights [N][C] split into tf32 hi / lo, box C x N
void launch_fused_dwpw_tc(const TmaMap& tm_in, const TmaMap& tm_out, const TmaMap& tm_whi, const TmaMap& tm_wlo, const TView& in,
                          const TView& out, const float* host_w, int stride, int dw_relu, int pw_relu, int frames,
                          cudaStream_t s) {
#define UF_T(CC, NN, SS, GG, BB) launch_fused_tc_t<CC, NN, SS, GG, BB>(tm_in, tm_out, tm_whi, tm_wlo, out, host_w, dw_relu, pw_relu, frames, s)
    const int C = in.C, N = out.C;
    // groups per CTA: as many tile pipelines as fit the SM's 227 KB (stride 2: the input slices alone take 74 KB each)
    // (groups, slice buffers) per shape, measured on B200: the group count matters most, a third slice buffer a little
    if (C == 16 && N == 32 && stride == 1) UF_T(16, 32, 1, 5, 2);
    else if (C == 32 && N == 32 && stride == 1) UF_T(32, 32, 1, 3, 3);
    else if (C == 32 && N == 32 && stride == 2) UF_T(32, 32, 2, 2, 2);
    else if (C == 32 && N == 64 && stride == 2) UF_T(32, 64, 2, 3, 1);
    else if (C == 32 && N == 64 && stride == 1) UF_T(32, 64, 1, 2, 4);
    else if (C == 64 && N == 64 && stride == 1) UF_T(64, 64, 1, 2, 2);
#undef UF_T
}

// ---------------------------------------------------------------------------------------------
// K6 on the tensor cores: dense 3x3 convolution (stride 1, pad = dilation, Cin <= 16, Cout <= 16: the RFB branches) as
// an implicit GEMM whose im2col costs nothing. A group stages the input tile (+ dilation halo) ONCE in shared memory
// in the canonical NO-SWIZZLE K-major operand layout, channel-chunk major: T[q][pixel][4 floats] (q = 4-channel chunk).
// In that layout 8 consecutive pixels of a tile row are one 8 x 16 B core matrix, the next 8-row group of the M tile
// (the next output row) is IW * 16 B further (SBO) and the next 16 bytes of K (the next chunk) NPIX * 16 B further (LBO)
// — and the operand of tap (ky, kx) is the SAME tile read from a start address shifted by ((ky*IW + kx) * dil) * 16 B.
// Nine taps x Cin/8 k-steps x 3 (3xTF32) MMAs of 128 x 16 x 8 per 8 x 16 pixel tile replace 9*Cin*Cout FFMAs per pixel;
// the SIMT work left is the tile staging (with the hi / lo split) and the 16-output epilogue. Same grouped layout as
// the fused kernel above: one CTA per SM, G independent tile pipelines.
// The swizzled form (UF_D3_SWZ, default) keeps the tile pixel-major, T[pixel][Cin floats] with the 32B / 64B swizzle
// applied from ABSOLUTE shared-memory address bits, SBO = one tile row; measured on B200: a descriptor whose start
// address is shifted by an arbitrary number of rows (not a multiple of the 8-row swizzle atom) reads such a tile
// correctly with base_offset = 0, i.e. the tensor core derives the swizzle phase from the address, not from the row index.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

struct Dense3Params {
    TView in, out;
    const float* w_hi;  // [9][CQ][16][4] floats: W[n][4q + j][tap], zero padded (n >= Cout, c >= Cin)
    const float* w_lo;
    int CQ;             // 4-channel chunks of K per tap, even (Cin padded to a multiple of 8)
    int dil, relu;
    int tiles_x, tiles_y, total_tiles;
    int group_bytes;    // 2 * CQ * NPIX * 16 rounded up to 128
    float bias[16];
};

template <int G, int U>  // G tile pipelines per CTA, U staging loads in flight per thread
__global__ void __launch_bounds__(128 * G, 1)
dense3x3_tc_kernel(const __grid_constant__ Dense3Params p) {
    constexpr int TX = 8, TY = 16;
    extern __shared__ __align__(128) uint8_t smem_d3[];
    const int tid = threadIdx.x, g = tid >> 7, lt = tid & 127, lwarp = lt >> 5;
    const int d = p.dil, IW = TX + 2 * d, IH = TY + 2 * d, NPIX = IH * IW, CQ = p.CQ;
    const int w_bytes = 9 * CQ * 256;
    uint8_t* w_hi = smem_d3;
    uint8_t* w_lo = w_hi + w_bytes;
#if UF_D3_SWZ
    // experiment: pixel-major tile [pix][CQ * 16 B] with the 32B / 64B swizzle keyed on ABSOLUTE address bits; group regions 1 KB aligned
    uint8_t* t_hi = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(w_lo + w_bytes) + 1023) & ~uintptr_t(1023)) + g * p.group_bytes;
    uint8_t* t_lo = t_hi + p.group_bytes / 2;
#else
    uint8_t* t_hi = w_lo + w_bytes + g * p.group_bytes;   // [CQ][NPIX][16 B]
    uint8_t* t_lo = t_hi + CQ * NPIX * 16;
#endif
#if UF_D3_SWZ
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(w_lo + w_bytes) + 1023) & ~uintptr_t(1023)) + G * p.group_bytes);
#else
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_lo + w_bytes + G * p.group_bytes);
#endif
    uint64_t* mma_bar = bars + g;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G);
    constexpr int NACC = 4;  // independent accumulators per group: consecutive MMAs into ONE accumulator serialise on its
                             // read-modify-write latency (~150 cycles each for these tiny N = 16 MMAs); the epilogue adds them up
    constexpr uint32_t TMEM_COLS = G * NACC * 16 <= 256 ? 256 : 512;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };

    pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < G; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weights: static, staged before the predecessor kernel is waited for
    for (int i = tid; i < w_bytes / 16; i += 128 * G) {
        reinterpret_cast<float4*>(w_hi)[i] = __ldg(reinterpret_cast<const float4*>(p.w_hi) + i);
        reinterpret_cast<float4*>(w_lo)[i] = __ldg(reinterpret_cast<const float4*>(p.w_lo) + i);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    const int vcta = blockIdx.x * G + g, vgrid = gridDim.x * G;
    const uint32_t acc = tmem_base + (uint32_t)(g * NACC * 16);
    const uint32_t taddr = acc + ((uint32_t)(lwarp * 32) << 16);
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const bool vec = ((p.out.C & 3) == 0) && ((p.out.pix_stride & 3) == 0) && ((reinterpret_cast<size_t>(p.out.p) & 15) == 0) &&
                     ((p.out.frame_stride & 3) == 0);
    const int C4 = p.in.C >> 2;  // real 4-channel chunks (<= CQ)
    int it = 0;
    for (int tile = vcta; tile < p.total_tiles; tile += vgrid, ++it) {
        int b = tile;
        const int txi = b % p.tiles_x; b /= p.tiles_x;
        const int tyi = b % p.tiles_y;
        const int f = b / p.tiles_y;
        const int x0 = txi * TX, y0 = tyi * TY;
        // stage the tile: T[q][pix] = in[y0 - d + py][x0 - d + px][4q..4q+3] (zero outside the map / past Cin), hi and lo
        const float* ip = p.in.p + (size_t)f * p.in.frame_stride;
        const int items = NPIX * CQ;
        for (int i0 = lt; i0 < items; i0 += U * 128) {
            float4 v[U];
            int so[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * 128;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                so[u] = -1;
                if (i < items) {
                    const int pix = i / CQ, q = i - pix * CQ;  // q fastest: a warp reads whole pixels (contiguous channels)
                    const int py = pix / IW, px = pix - py * IW;
                    const int gy = y0 - d + py, gx = x0 - d + px;
#if UF_D3_SWZ
                    {
                        const uint32_t rowa = smem_u32(t_hi) + (uint32_t)pix * (uint32_t)(CQ * 16);
                        const int sw = CQ == 4 ? (int)((rowa >> 7) & 3u) : (int)((rowa >> 7) & 1u);
                        so[u] = pix * (CQ * 16) + ((q ^ sw) << 4);
                    }
#else
                    so[u] = (q * NPIX + pix) * 16;
#endif
                    if (q < C4 && gy >= 0 && gy < p.in.H && gx >= 0 && gx < p.in.W)
                        v[u] = *reinterpret_cast<const float4*>(ip + ((size_t)gy * p.in.W + gx) * p.in.pix_stride + q * 4);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (so[u] >= 0) {
                    float4 l;
                    l.x = tf32_lo(v[u].x);
                    l.y = tf32_lo(v[u].y);
                    l.z = tf32_lo(v[u].z);
                    l.w = tf32_lo(v[u].w);
                    *reinterpret_cast<float4*>(t_hi + so[u]) = v[u];  // kind::tf32 ignores the low 13 mantissa bits: raw = hi
                    *reinterpret_cast<float4*>(t_lo + so[u]) = l;
                }
            }
        }
        fence_proxy_async();
        group_sync();
        if (lt == 0) {
            tc_fence_after();
            const uint32_t lbo_a = (uint32_t)NPIX * 16u, sbo_a = (uint32_t)IW * 16u;
            int nmma = 0;
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - ky * 3;
                const uint32_t shift = (uint32_t)((ky * d * IW + kx * d) * 16);
                for (int ks = 0; ks < CQ / 2; ++ks) {
#if UF_D3_SWZ
                    const uint32_t rowb = (uint32_t)CQ * 16u;
                    const uint32_t a_off = (uint32_t)((ky * d * IW + kx * d)) * rowb + (uint32_t)ks * 32u;
                    const uint64_t lay = CQ == 4 ? 4ull : 6ull;  // SWIZZLE_64B / SWIZZLE_32B
                    const uint32_t a_hi_addr = smem_u32(t_hi) + a_off, a_lo_addr = smem_u32(t_lo) + a_off;
                    const uint64_t d_ahi = (uint64_t)((a_hi_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(((uint32_t)IW * rowb) >> 4) << 32) | (1ull << 46) | (lay << 61);
                    const uint64_t d_alo = (uint64_t)((a_lo_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(((uint32_t)IW * rowb) >> 4) << 32) | (1ull << 46) | (lay << 61);
#else
                    const uint32_t a_off = (uint32_t)(2 * ks) * lbo_a + shift;
                    const uint64_t d_ahi = umma_desc_nosw(smem_u32(t_hi) + a_off, lbo_a, sbo_a);
                    const uint64_t d_alo = umma_desc_nosw(smem_u32(t_lo) + a_off, lbo_a, sbo_a);
#endif
                    const uint32_t w_off = (uint32_t)((tap * CQ + 2 * ks) * 256);
                    const uint64_t d_whi = umma_desc_nosw(smem_u32(w_hi) + w_off, 256u, 128u);
                    const uint64_t d_wlo = umma_desc_nosw(smem_u32(w_lo) + w_off, 256u, 128u);
                    umma_tf32(acc + 16 * (nmma % NACC), d_alo, d_whi, idesc, nmma >= NACC ? 1u : 0u); ++nmma;
                    umma_tf32(acc + 16 * (nmma % NACC), d_ahi, d_wlo, idesc, nmma >= NACC ? 1u : 0u); ++nmma;
                    umma_tf32(acc + 16 * (nmma % NACC), d_ahi, d_whi, idesc, nmma >= NACC ? 1u : 0u); ++nmma;
                }
            }
            umma_commit(mma_bar);
        }
        __syncwarp();
        mbar_wait(mma_bar, (uint32_t)(it & 1));
        tc_fence_after();
        float v[16];
        {
            float part[32];
            tmem_ld32(taddr, part, true);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = part[j] + part[16 + j];
            tmem_ld32(taddr + 32, part, true);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += part[j] + part[16 + j];
        }
        tc_fence_before();
        const int ox = x0 + (lt & 7), oy = y0 + (lt >> 3);  // M row = ty * 8 + tx
        if (ox < p.out.W && oy < p.out.H) {
            float* op = p.out.p + (size_t)f * p.out.frame_stride + ((size_t)oy * p.out.W + ox) * p.out.pix_stride;
            if (vec) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    if (j < p.out.C) {
                        float4 x = make_float4(v[j] + p.bias[j], v[j + 1] + p.bias[j + 1], v[j + 2] + p.bias[j + 2], v[j + 3] + p.bias[j + 3]);
                        if (p.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                        *reinterpret_cast<float4*>(op + j) = x;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (j < p.out.C) {
                        const float x = v[j] + p.bias[j];
                        op[j] = p.relu ? fmaxf(x, 0.f) : x;
                    }
                }
            }
        }
        group_sync();  // every thread has read its accumulator row and the MMAs are done with the tile: both may be overwritten
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

bool dense3x3_tc_supported(int cin, int cout, int dil) {
    return cin % 4 == 0 && cin >= 4 && cin <= 16 && cout >= 1 && cout <= 16 && dil >= 1 && dil <= 8;
}

size_t dense3x3_tc_weight_floats(int cin) { return (size_t)9 * (((cin + 7) / 8 * 8) / 4) * 16 * 4; }

// w_hi / w_lo: DEVICE arrays in the layout of Dense3Params; host_bias: Cout floats (HOST)
void launch_dense3x3_tc(const TView& in, const TView& out, const float* w_hi, const float* w_lo, const float* host_bias, int dil,
                        int relu, int frames, cudaStream_t s) {
    Dense3Params p;
    p.in = in; p.out = out; p.w_hi = w_hi; p.w_lo = w_lo;
    p.CQ = ((in.C + 7) / 8 * 8) / 4;
    p.dil = dil; p.relu = relu;
    p.tiles_x = (out.W + 7) / 8; p.tiles_y = (out.H + 15) / 16;
    p.total_tiles = p.tiles_x * p.tiles_y * frames;
    const int npix = (16 + 2 * dil) * (8 + 2 * dil);
#if UF_D3_SWZ
    p.group_bytes = (2 * p.CQ * npix * 16 + 2047) / 2048 * 2048;
#else
    p.group_bytes = (2 * p.CQ * npix * 16 + 127) / 128 * 128;
#endif
    for (int i = 0; i < 16; ++i) p.bias[i] = i < out.C ? host_bias[i] : 0.f;
    // as many tile pipelines as fit ~200 KB next to the weights (a TMEM kernel gets one CTA per SM)
    const size_t fixed = (size_t)2 * 9 * p.CQ * 256 + 8 * 8 + 16 + 128 + 1024;
    const int fit = (int)((200 * 1024 - fixed) / p.group_bytes);
    const int G = fit >= 6 ? 6 : (fit >= 4 ? 4 : 3);
    const size_t smem = fixed + (size_t)G * p.group_bytes;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(dense3x3_tc_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(dense3x3_tc_kernel<4, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(dense3x3_tc_kernel<6, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured[dev & 63] = true;
    }
    int grid = 148;
    if (grid > (p.total_tiles + G - 1) / G) grid = (p.total_tiles + G - 1) / G;
    if (G == 6) launch_pdl(dense3x3_tc_kernel<6, 8>, dim3(grid), dim3(128 * 6), smem, s, p);
    else if (G == 4) launch_pdl(dense3x3_tc_kernel<4, 12>, dim3(grid), dim3(128 * 4), smem, s, p);
    else launch_pdl(dense3x3_tc_kernel<3, 16>, dim3(grid), dim3(128 * 3), smem, s, p);
}

#ifdef UF_TC_TIMING
void tc_timing_read(long long* out16) { cudaMemcpyFromSymbol(out16, g_tc_timing, sizeof(long long) * 16); }
#endif

static bool g_pdl = false;  // measured: no gain once the chain is replayed as a CUDA graph, so opt-in (UF_FLAG_PDL)
bool pdl_enabled() { return g_pdl; }
void pdl_set_enabled(bool on) { g_pdl = on; }

bool pointwise_tc_supported(int K, int N) { return K >= 32 && K % 4 == 0 && N >= 1 && N <= 256; }

int pointwise_tc_n_umma(int N) { return (N + 15) / 16 * 16; }

// fp32 matrix [rows][cols] (row pitch row_stride_bytes) for the epilogue's TMA store: box 32 x 32, 128B swizzle
bool make_tmap_f32_2d_store(TmaMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim,
                    gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// host_bias: N floats in HOST memory (copied into the kernel parameter space); tm_out may be null (legacy epilogue)
void launch_pointwise_tc(const TmaMap& tm_a, const TmaMap& tm_whi, const TmaMap& tm_wlo, const TmaMap* tm_out,
                         const TmaMap* tm_res, const TView& in, const TView& out, const TView* res, const float* host_bias,
                         int relu, int frames, cudaStream_t s, const TcDepthwise* dw) {
    TcParams p;
    p.out = out;
    p.res = res ? *res : TView{};
    p.has_res = res != nullptr;
    p.relu = relu;
    p.M = frames * in.H * in.W;
    p.K = in.C;
    p.N = out.C;
    p.n_umma = pointwise_tc_n_umma(p.N);
    p.tma_store = (tm_out != nullptr && (res == nullptr || tm_res != nullptr)) ? 1 : 0;
    p.dw_mode = dw ? 1 : 0;
    if (dw) {
        p.dw_stride = dw->stride; p.dw_relu = dw->relu; p.dw_in = dw->in; p.dw_w = dw->w; p.dw_b = dw->b;
        p.M = frames * out.H * out.W;  // the 1x1 conv runs on the depthwise OUTPUT pixels
        p.K = dw->in.C;
    } else {
        p.dw_stride = 1; p.dw_relu = 0; p.dw_in = TView{}; p.dw_w = nullptr; p.dw_b = nullptr;
    }
    for (int i = 0; i < 256; ++i) p.bias[i] = i < p.N ? host_bias[i] : 0.f;
    const int stage_bytes = 2 * TC_A_BYTES + 2 * p.n_umma * TC_BK * 4;
    int stages = (188 * 1024) / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    p.stages = stages;
    p.wide = p.n_umma <= 128 ? 1 : 0;  // 2 * n_umma <= 256 (the MMA's N limit) and two such accumulators fit 512 columns
    int cols = 32;
    while (cols < 2 * (p.wide ? 2 * p.n_umma : p.n_umma)) cols <<= 1;
    p.tmem_cols = cols;
    // stages | barriers + tmem slot | (pad to 1 KB) | 4 per-warp staging tiles of 4.6 KB (legacy) / 4 KB (TMA store)
    const size_t smem = (size_t)stages * stage_bytes + (3 * stages + 12) * sizeof(uint64_t) + 16 + 1024 + 8 * 4096 + 1024;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured[dev & 63] = true;
    }
    const int tiles = (p.M + TC_BM - 1) / TC_BM;
    const int grid = tiles < 148 ? tiles : 148;
    const TmaMap& to = tm_out ? *tm_out : tm_a;  // unused when tma_store == 0
    const TmaMap& tr = tm_res ? *tm_res : tm_a;  // unused without a residual
    launch_pdl(pw_tc_kernel, dim3(grid), dim3(TC_THREADS), smem, s, *reinterpret_cast<const CUtensorMap*>(&tm_a),
               *reinterpret_cast<const CUtensorMap*>(&tm_whi), *reinterpret_cast<const CUtensorMap*>(&tm_wlo),
               *reinterpret_cast<const CUtensorMap*>(&to), *reinterpret_cast<const CUtensorMap*>(&tr), p);
}

}  // namespace uf
