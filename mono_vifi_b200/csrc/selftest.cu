// Device self-tests exported through the C ABI (used by tests/, not by the product path).
#include "../../include/monovifi_b200.h"
#include "common.cuh"

namespace {
__global__ void div_selftest_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ q_seq,
                               float* __restrict__ q_ieee, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_seq[i] = mvf::div_with(a[i], b[i], mvf::rcp_refined(b[i]));
    q_ieee[i] = __fdiv_rn(a[i], b[i]);
}
}  // namespace

extern "C" int mvf_selftest_division(const float* a, const float* b, float* q_sequence, float* q_ieee, size_t n,
                                     void* stream) {
    if (!a || !b || !q_sequence || !q_ieee || n == 0) return MVF_ERR_INVALID;
    div_selftest_k<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, q_sequence, q_ieee, n);
    return cudaGetLastError() == cudaSuccess ? MVF_OK : MVF_ERR_CUDA;
}
