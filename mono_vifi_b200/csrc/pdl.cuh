// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// be scheduled while its predecessor in the stream is still draining; its CTAs run their prologue (barrier init, TMEM
// allocation, descriptor prefetch) and then block in griddepcontrol.wait until the predecessor has COMPLETED and its
// memory is visible.  The step is ~560 short launches (10-60 us each) replayed from one CUDA graph, so the launch /
// drain / prologue gap between two kernels is a measurable share of it.
//
// Contract for every kernel launched through launch_pdl():
//   * every thread executes pdl_wait() before its first access to global memory (reads AND writes: the predecessor may
//     still be reading a buffer this kernel overwrites);  transitivity (kernel N+1 waits for N, which waited for N-1)
//     relies on the wait being unconditional;
//   * pdl_trigger() right after it lets the successor be scheduled as soon as SM resources free up.
// Kernels launched without the attribute (torch's, F1, the optimiser) keep full stream serialisation on both sides; the
// two instructions are no-ops there.  MVF_PDL=0 disables the attribute (A/B timing, debugging).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace mvf {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
    pdl_wait();
    pdl_trigger();
}

inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("MVF_PDL");
        v = e ? (std::atoi(e) != 0) : 1;
    }
    return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace mvf
