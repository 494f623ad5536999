// Channels-last data-movement kernels of the depth decoder (see ops_cl.cu).
#pragma once
#include <cuda_runtime.h>

namespace mvf {
cudaError_t upcat_pad_fwd(const float* a, const float* skip, float* y, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st);
cudaError_t upcat_pad_bwd(const float* gy, float* ga, float* gskip, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st);
}  // namespace mvf
