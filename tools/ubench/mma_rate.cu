// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (K=8) vs kind::f16 (bf16, K=16), M=128, cta_group::1.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) { return (uint64_t)((a >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }
template <int KIND>  // 0 tf32, 1 bf16
__global__ void bench(int n_umma, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("fence.proxy.async.shared::cta;");
    uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n_umma >> 3) << 17) | (8u << 24);
        uint64_t ad = desc_sw128(smem_u32(smem)), bd = desc_sw128(smem_u32(smem + 16384));
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t acc = i > 0;
            uint64_t a2 = ad + ((i & 3) * 2), b2 = bd + ((i & 3) * 2);
            if (KIND == 0)
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tm), "l"(a2), "l"(b2), "r"(idesc), "r"(acc));
            else
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tm), "l"(a2), "l"(b2), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)));
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int grid : {1, 148})
        for (int n : {64, 128, 256}) {
            for (int kind = 0; kind < 2; ++kind) {
                const int iters = 2000;
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    if (kind == 0) bench<0><<<grid, 128, 50 * 1024>>>(n, iters, d); else bench<1><<<grid, 128, 50 * 1024>>>(n, iters, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                }
                const double cyc = (double)h / iters;
                const int K = kind == 0 ? 8 : 16;
                printf("grid %3d N %3d %s: %.1f cycles/MMA, %.0f MAC/clk/SM\n", grid, n, kind == 0 ? "tf32 K=8 " : "bf16 K=16", cyc, 128.0 * n * K / cyc);
            }
        }
    return 0;
}
