// Lite-Mono block kernels (see litemono.cu): dense channels-last [P pixels][C], C % 4 == 0.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mvf {
// w_taps: [9][C] (tap-major copy of the [C,1,3,3] filter); flip = 1: data gradient (mirrored taps)
cudaError_t dwconv3x3_fwd(const float* x, const float* w_taps, const float* bias, float* y, int B, int C, int H, int W, int dil, int flip,
                          cudaStream_t st);
size_t dwconv3x3_wgrad_workspace_floats(long long P, int C);
cudaError_t dwconv3x3_wgrad(const float* x, const float* gy, float* gw, float* workspace, int B, int C, int H, int W, int dil, cudaStream_t st);
cudaError_t gelu_fwd(const float* x, float* y, long long n, cudaStream_t st);
cudaError_t gelu_bwd(const float* x, const float* gy, float* gx, long long n, cudaStream_t st);
cudaError_t layernorm_cl_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, long long P, int C, float eps,
                             cudaStream_t st);
size_t layernorm_bwd_workspace_floats(long long P, int C);
cudaError_t layernorm_cl_bwd(const float* x, const float* gy, const float* w, const float* mean, const float* rstd, float* gx, float* gw,
                             float* gb, float* workspace, long long P, int C, cudaStream_t st);
}  // namespace mvf
