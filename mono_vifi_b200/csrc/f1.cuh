// Argument blocks of the fused view-synthesis + photometric-loss kernels (F1), shared by fwd / bwd / api.
#pragma once
#include "common.cuh"

namespace mvf {

// Cross-CTA accumulators: integer fixed-point so the result does not depend on CTA scheduling order.
// The last CTA to finish reads them, writes the float outputs and zeroes them again (self-cleaning), so
// the workspace only has to be zeroed once, when it is created (mvf_workspace_init).
struct F1Workspace {
    unsigned int counter_fwd;
    unsigned int counter_bwd;
    unsigned int pad[2];
    long long acc[1];  // fwd: [B][4] {photo, Sx, Sy, sum_disp}; bwd: [2][B][12] grad_P   (B*28 entries in total)
};
__host__ __device__ inline size_t f1_workspace_bytes(int B) { return 16 + sizeof(long long) * (size_t)B * 28; }
__host__ __device__ inline long long* ws_fwd_acc(F1Workspace* w) { return w->acc; }
__host__ __device__ inline long long* ws_bwd_acc(F1Workspace* w, int B) { return w->acc + 4 * (size_t)B; }

struct F1Args {
    // inputs (device pointers, contiguous fp32 NCHW)
    const float* disp;   // [B,1,H,W]
    const float* tgt;    // [B,3,H,W]
    const float* src0;   // [B,3,H,W]
    const float* src1;   // [B,3,H,W]
    const float* inv_K;  // [B,4,4]
    const float* P0;     // [B,3,4]  (K@T0)[:, :3]
    const float* P1;     // [B,3,4]
    const float* noise;  // [B,nid,H,W] or null
    const float* mask;   // [B,1,H,W] or null
    // forward outputs
    float* loss;         // [4] {total, photometric mean, smoothness (unweighted), 0}
    float* stats;        // [B,4] {mean(disp), Sx, Sy, 0}  (saved for backward)
    uint8_t* idx;        // [B,H,W] argmin channel of `combined` (train.py:1033)
    // optional forward debug outputs (null in production)
    int* x0y0;           // [2 src][2 (x,y)][B,H,W]
    float* warp0;        // [B,3,H,W]
    float* warp1;
    float* to_opt;       // [B,H,W]
    // backward
    const float* gout;   // device scalar dL/dloss (null = 1)
    float* g_disp;       // [B,1,H,W]
    float* g_P0;         // [B,3,4]
    float* g_P1;
    F1Workspace* ws;
    int B, H, W;
    float min_disp, disp_range, smooth_w;
    int flags;
};

cudaError_t launch_f1_forward(const F1Args& a, cudaStream_t stream);
// persistent TMA / warp-specialised forward (f1_fwd_tma.cu); launch_f1_forward dispatches to it when the shape qualifies
bool f1_forward_tma_eligible(const F1Args& a);
cudaError_t launch_f1_forward_tma(const F1Args& a, cudaStream_t stream);
cudaError_t launch_f1_backward(const F1Args& a, cudaStream_t stream);

}  // namespace mvf
