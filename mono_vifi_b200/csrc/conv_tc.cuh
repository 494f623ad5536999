// Tensor-core (tcgen05 / TMEM / TMA) convolution kernels: host-side entry points.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mvf {
namespace tc {

// convolution problem on channels-last tensors.  Input x[b][y][x][c] with element strides (x_sB, x_sH, x_sW, 1);
// output y[b][oy][ox][co] with strides (y_sB, y_sH, y_sW, 1), Ho = (H + 2 pad - KH) / stride + 1.
struct ConvDesc {
    int B, Cin, H, W;
    int Cout, KH, KW, pad, stride, stride_x;  // stride: rows, stride_x: columns
    long long x_sB, x_sH, x_sW;
    long long y_sB, y_sH, y_sW;
};

// wgrad problem on channels-last tensors: dw[Cout,Cin,KH,KW] = sum_pixels gy . x  (conv_wgrad.cu)
struct WgradDesc {
    int B, Cin, H, W;
    int Cout, KH, KW, pad, stride, stride_x;
    long long x_sB, x_sH, x_sW;
    long long g_sB, g_sH, g_sW;
};
const char* wgrad_check(const WgradDesc& d);
size_t wgrad_workspace_floats(const WgradDesc& d);
cudaError_t conv_wgrad(const WgradDesc& d, const float* x, const float* gy, float* dw, float* workspace, cudaStream_t st,
                       const char** why);

size_t packed_filter_floats(int N, int K, int KH, int KW);
cudaError_t pack_filters(const float* w, float* out, int Cout, int Cin, int KH, int KW, int dgrad, cudaStream_t st);
// all filter banks of a step in one launch; table (device): n_entries x {w, out, Cout, Cin, KH, KW, dgrad, first block} as int64,
// every bank takes ceil(packed floats / pack_chunk()) blocks
int pack_chunk();
cudaError_t pack_filters_multi(const long long* table, int n_entries, long long total_blocks, cudaStream_t st);
cudaError_t umma_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn_major, cudaStream_t st,
                          int row_off = 0, int base_off_mode = 0);
void set_debug_buffer(float* p);  // development aid: dump of pipeline stage 0, see conv_tc.cu
const char* conv_check(const ConvDesc& d);  // nullptr if the tcgen05 path covers the problem, else the reason
// act: 0 none, 1 relu, 2 elu, 3 prelu with per-channel `slope`
cudaError_t conv_forward(const ConvDesc& d, const float* x, const float* w_packed, const float* bias, float* y, int act,
                         cudaStream_t st, const char** why, const float* slope = nullptr);
// data gradient of a stride-2 convolution (d = the forward problem: x_* describe gx, y_* describe gy); see conv_tc.cu
const char* conv_dgrad_s2_check(const ConvDesc& d);
// (with `bias` it is the forward of nn.ConvTranspose2d(stride 2): the same arithmetic with the bias added in the epilogue)
cudaError_t conv_dgrad_s2(const ConvDesc& d, const float* gy, const float* w_packed, float* gx, cudaStream_t st, const char** why,
                          const float* bias = nullptr);
int conv_dgrad_s2_plan_table(int KH, int KW, int pad, int* out, int capacity);

}  // namespace tc
}  // namespace mvf
