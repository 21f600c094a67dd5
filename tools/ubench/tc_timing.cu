// Role timing of pw_tc_kernel (built with -DUF_TC_TIMING): total vs waiting cycles per warp role for one layer shape.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../../infercam_onnx_b200/csrc/kernels.h"
namespace uf { void tc_timing_read(long long*); }
using namespace uf;
int main(int argc, char** argv) {
    const int frames = 128, H = argc > 1 ? atoi(argv[1]) : 30, W = argc > 2 ? atoi(argv[2]) : 40;
    const int K = argc > 3 ? atoi(argv[3]) : 64, N = argc > 4 ? atoi(argv[4]) : 64;
    const long long M = (long long)frames * H * W;
    float *a, *o, *whi, *wlo;
    cudaMalloc(&a, M * K * 4); cudaMalloc(&o, M * N * 4); cudaMalloc(&whi, N * K * 4); cudaMalloc(&wlo, N * K * 4);
    cudaMemset(a, 0, M * K * 4); cudaMemset(whi, 0, N * K * 4); cudaMemset(wlo, 0, N * K * 4);
    TmaMap ta, th, tl, to;
    make_tmap_f32_2d(&ta, a, M, K, K * 4, 128);
    make_tmap_f32_2d(&th, whi, N, K, K * 4, pointwise_tc_n_umma(N));
    make_tmap_f32_2d(&tl, wlo, N, K, K * 4, pointwise_tc_n_umma(N));
    make_tmap_f32_2d_store(&to, o, M, N, N * 4);
    TView in{a, (long long)H * W * K, K, K, H, W}, out{o, (long long)H * W * N, N, N, H, W};
    std::vector<float> bias(N, 0.f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch_pointwise_tc(ta, th, tl, &to, nullptr, in, out, nullptr, bias.data(), 1, frames, 0);
    cudaEventRecord(e0);
    launch_pointwise_tc(ta, th, tl, &to, nullptr, in, out, nullptr, bias.data(), 1, frames, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long t[16]; tc_timing_read(t);
    const char* names[4] = {"producer", "mma", "converter", "epilogue"};
    printf("%dx%d K=%d N=%d M=%lld tiles=%lld: %.1f us (%s)\n", H, W, K, N, M, (M + 127) / 128, ms * 1e3, cudaGetErrorString(e));
    for (int r = 0; r < 4; ++r) printf("  %-9s total %8lld cycles, waiting %8lld (%.0f%%)\n", names[r], t[r * 2], t[r * 2 + 1], 100.0 * t[r * 2 + 1] / (t[r * 2] + 1));
    printf("  prologue %lld cycles, entry->teardown barrier %lld cycles (block 0)\n", t[8], t[9]);
    return 0;
}
