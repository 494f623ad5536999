// Stand-alone (unfused) kernels behind the reference's layers.py module API.
#pragma once
#include "common.cuh"

namespace mvf {

cudaError_t disp_to_depth_fwd(const float* disp, float* sd, float* depth, size_t n, float min_disp, float range,
                              cudaStream_t st);
cudaError_t disp_to_depth_bwd(const float* disp, const float* g_sd, const float* g_depth, float* g_disp, size_t n,
                              float min_disp, float range, cudaStream_t st);
cudaError_t backproject_fwd(const float* depth, const float* inv_K, float* out, int B, int H, int W, cudaStream_t st);
cudaError_t backproject_bwd(const float* g_out, const float* inv_K, float* g_depth, int B, int H, int W,
                            cudaStream_t st);
cudaError_t project_fwd(const float* points, const float* P, float* grid_out, int B, int H, int W, float eps,
                        cudaStream_t st);
cudaError_t project_bwd(const float* points, const float* P, const float* g_grid, float* g_points, float* g_P,
                        long long* acc, int B, int H, int W, float eps, cudaStream_t st);
cudaError_t ssim_fwd(const float* x, const float* y, float* out, int N, int H, int W, cudaStream_t st);
cudaError_t ssim_bwd(const float* x, const float* y, const float* g_out, float* g_x, float* coef_ws, int N, int H,
                     int W, cudaStream_t st);
cudaError_t smooth_fwd(const float* disp, const float* img, float* out, long long* acc, int B, int H, int W,
                       cudaStream_t st);
cudaError_t smooth_bwd(const float* disp, const float* img, const float* gout, float* g_disp, int B, int H, int W,
                       cudaStream_t st);
cudaError_t si_log_fwd(const float* pred, const float* target, const float* mask, float* loss, float* stats,
                       long long* acc, int B, size_t HW, float beta, cudaStream_t st);
cudaError_t si_log_bwd(const float* pred, const float* target, const float* mask, const float* stats,
                       const float* gout, float* g_pred, float* g_target, int B, size_t HW, float beta,
                       cudaStream_t st);

}  // namespace mvf
