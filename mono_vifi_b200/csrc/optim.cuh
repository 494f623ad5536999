// Fused gradient clipping + AdamW over a flat parameter arena (see optim.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mvf {
size_t adamw_workspace_bytes();
cudaError_t adamw_step(float* p, const float* g, float* m, float* v, long long n, float* state, void* workspace, float lr,
                       float beta1, float beta2, float eps, float wd, float max_norm, cudaStream_t st);
}  // namespace mvf
