#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k_empty(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_tmem(int* p) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t t = slot;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(t));
    if (p && threadIdx.x == 9999) *p = 1;
}
template <typename F> float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 5; ++i) f();
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 20; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best * 1e3f;
}
int main() {
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    printf("empty <<<148,448,0>>>      %.1f us\n", time_it([] { k_empty<<<148, 448, 0>>>(nullptr); }));
    printf("empty <<<148,448,220KB>>>  %.1f us\n", time_it([] { k_empty<<<148, 448, 220 * 1024>>>(nullptr); }));
    printf("empty <<<148,256,48KB>>>   %.1f us\n", time_it([] { k_empty<<<148, 256, 48 * 1024>>>(nullptr); }));
    printf("tmem  <<<148,448,220KB>>>  %.1f us\n", time_it([] { k_tmem<<<148, 448, 220 * 1024>>>(nullptr); }));
    printf("tmem  <<<148,448,0>>>      %.1f us\n", time_it([] { k_tmem<<<148, 448, 0>>>(nullptr); }));
    printf("alternating smem configs    %.1f us per pair\n", time_it([] { k_empty<<<148, 448, 220 * 1024>>>(nullptr); k_empty<<<148, 256, 0>>>(nullptr); }));
    return 0;
}
