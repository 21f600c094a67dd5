// pdl.cuh — programmatic dependent launch (PDL). The kernel chain of a stage is ~49 short launches; with PDL a
// kernel is launched while its predecessor is still running, does its private prologue (barrier init, TMEM
// allocation, weights into shared memory) and only then waits for the predecessor's results, so launch latency
// and prologues hide behind the previous kernel's tail.
//   device: pdl_launch_dependents() first thing, pdl_wait() before the first access to data produced upstream
//           (also before the first global write: the predecessor may still be reading what we overwrite... no
//           buffer is reused inside a stage, but the wait keeps the rule simple). Both are no-ops in plain launches.
//   host:   launch_pdl(kernel, grid, block, smem, stream, args...) adds the stream-serialization attribute.
#pragma once
#include <cuda_runtime.h>

namespace uf {

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();
void pdl_set_enabled(bool on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace uf
