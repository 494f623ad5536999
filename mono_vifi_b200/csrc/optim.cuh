// Fused gradient clipping + AdamW over a flat parameter arena (see optim.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mvf {
size_t adamw_workspace_bytes();
cudaError_t adamw_step(float* p, const float* g, float* m, float* v, long long n, long long n_dup, const unsigned char* skip,
                       float* state, void* workspace, float beta1, float beta2, float eps, float wd, float max_norm,
                       cudaStream_t st);
// copies n_tensors gradient tensors (host arrays of device pointers / arena offsets / element counts) into the arena;
// a null source zero-fills its slice
cudaError_t gather_grads(float* G, const void* const* srcs, const long long* offsets, const long long* sizes, int n_tensors,
                         cudaStream_t st);
}  // namespace mvf
