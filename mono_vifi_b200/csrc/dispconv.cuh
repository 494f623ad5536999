// 3x3 convolution with one output channel on a padded channels-last input (see dispconv.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mvf {
cudaError_t dispconv_fwd(const float* xp, const float* w, const float* bias, float* y, int B, int C, int H, int W, cudaStream_t st);
cudaError_t dispconv_dgrad(const float* gy, const float* w, float* gxp, int B, int C, int H, int W, cudaStream_t st);
size_t dispconv_wgrad_workspace_floats(long long P, int C);
cudaError_t dispconv_wgrad(const float* xp, const float* gy, float* gw, float* gb, float* workspace, int B, int C, int H, int W, cudaStream_t st);
}  // namespace mvf
