// Fused training-mode BatchNorm2d (+ residual add) + ReLU on dense channels-last tensors (see bn_cl.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mvf {
size_t bn_workspace_floats(long long P, int C);
cudaError_t bn_forward(const float* x, const float* identity, float* y, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, long long* num_batches_tracked, float* save_mean, float* save_invstd, float* workspace,
                       long long P, int C, float eps, float momentum, int relu, cudaStream_t st);
cudaError_t bn_backward(const float* x, const float* gy, const float* y, const float* gamma, const float* save_mean,
                        const float* save_invstd, float* gx, float* gidentity, float* dgamma, float* dbeta, float* workspace,
                        long long P, int C, int relu, cudaStream_t st);
// SyncBatchNorm (train.py:205-208): the same kernels split around a cross-rank sum of `sums` ([2C + 1] doubles:
// two per-channel sums and the pixel count).  *_stats_* produce this rank's sums, *_apply_* consume the global ones.
cudaError_t bn_sync_stats_fwd(const float* x, double* sums, float* workspace, long long P, int C, cudaStream_t st);
cudaError_t bn_sync_apply_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, long long* num_batches_tracked, float* save_mean,
                              float* save_invstd, const double* sums, long long P, int C, float eps, float momentum, int relu,
                              cudaStream_t st);
cudaError_t bn_sync_stats_bwd(const float* x, const float* gy, const float* y, const float* save_mean, const float* save_invstd,
                              double* sums, float* dgamma, float* dbeta, float* workspace, long long P, int C, int relu,
                              cudaStream_t st);
cudaError_t bn_sync_apply_bwd(const float* x, const float* gy, const float* y, const float* gamma, const float* save_mean,
                              const float* save_invstd, float* gx, float* gidentity, const double* sums, float* scratch,
                              long long P, int C, int relu, cudaStream_t st);
// gpre = gy * act'(y) (act 0 none / 1 relu / 2 elu, from the saved output; gpre may be null for act 0) and, when gbias is
// given, gbias[c] = sum over pixels of gpre; workspace: bn_workspace_floats(P, C)
cudaError_t act_bwd_bias(const float* gy, const float* y, float* gpre, float* gbias, float* workspace, long long P, int C, int act,
                         cudaStream_t st);
}  // namespace mvf
