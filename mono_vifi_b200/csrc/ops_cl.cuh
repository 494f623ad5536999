// Channels-last data-movement kernels of the depth decoder (see ops_cl.cu).
#pragma once
#include <cuda_runtime.h>

namespace mvf {
cudaError_t upcat_pad_fwd(const float* a, const float* skip, float* y, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st);
cudaError_t upcat_pad_bwd(const float* gy, float* ga, float* gskip, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st);
// MaxPool2d(3, stride 2, padding 1) on dense channels-last tensors; idx holds the in-window position (0..8) of each maximum
cudaError_t maxpool3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int C, int H, int W, cudaStream_t st);
cudaError_t maxpool3s2_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int C, int H, int W, cudaStream_t st);
}  // namespace mvf
