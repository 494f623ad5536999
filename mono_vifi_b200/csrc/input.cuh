// GPU input pipeline (see input.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mvf {
size_t input_pipeline_workspace_floats(int B, int F);
cudaError_t input_pipeline(const unsigned char* frames, const float* prm_f, const int* prm_i, float* workspace, float* const* color_dev,
                           float* const* color_aug_dev, int B, int F, int H, int W, cudaStream_t st);
}  // namespace mvf
