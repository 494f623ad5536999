// Evaluation-path kernels (see eval.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mvf {
cudaError_t bn_eval_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, long long P, int C, float eps, int relu, cudaStream_t st);
size_t depth_eval_workspace_bytes(int Hg, int Wg);
cudaError_t depth_eval(const float* disp, int h, int w, const float* gt, int Hg, int Wg, float min_d, float max_d, int eigen_crop,
                       float stereo_scale, void* workspace, float* metrics8, cudaStream_t st);
}  // namespace mvf
