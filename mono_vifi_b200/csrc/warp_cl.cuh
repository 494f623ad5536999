// Feature warp ("F2") and bilinear resize kernels (see warp_cl.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mvf {
// layout 0: dense NCHW (any C); layout 1: dense channels-last, C % 4 == 0 (element (b,c,y,x) at ((b*H + y)*W + x)*C + c)
cudaError_t resize_bilinear_fwd(const float* x, float* y, int B, int C, int Hi, int Wi, int Ho, int Wo, float scale_h, float scale_w,
                                int align_corners, float mul_even, float mul_odd, int layout, cudaStream_t st);
cudaError_t resize_bilinear_bwd(const float* gy, float* gx, int B, int C, int Hi, int Wi, int Ho, int Wo, float scale_h, float scale_w,
                                int align_corners, float mul_even, float mul_odd, int layout, cudaStream_t st);
// backward warp by a pixel-unit flow [B,2,H,W] (NCHW), border padding, align_corners=True
cudaError_t flow_warp_fwd(const float* x, const float* flow, float* y, int B, int C, int H, int W, int layout, cudaStream_t st);
size_t flow_warp_bwd_workspace_bytes(int B, int C, int H, int W);
cudaError_t flow_warp_bwd(const float* gy, const float* flow, float* gx, int B, int C, int H, int W, void* workspace, size_t workspace_bytes,
                          cudaStream_t st);
// y = prelu(x + res) with one slope per channel; channels-last, C % 4 == 0; res may be null
cudaError_t prelu_cl_fwd(const float* x, const float* res, const float* slope, float* y, long long P, int C, cudaStream_t st);
// Rodrigues + translation -> 4x4 (layers.py:28-103); one thread per batch item
cudaError_t pose_matrix_fwd(const float* axisangle, const float* translation, float* M, int B, int invert, cudaStream_t st);
cudaError_t pose_matrix_bwd(const float* axisangle, const float* translation, const float* gM, float* g_axisangle, float* g_translation,
                            int B, int invert, cudaStream_t st);
}  // namespace mvf
