// C ABI of libmonovifi_b200.so (include/monovifi_b200.h).  Plain pointers and sizes only.
#include "../../include/monovifi_b200.h"

#include <cstdio>
#include <cstring>

#include "bn_cl.cuh"
#include "conv_tc.cuh"
#include "dispconv.cuh"
#include "eval.cuh"
#include "f1.cuh"
#include "input.cuh"
#include "litemono.cuh"
#include "ops.cuh"
#include "ops_cl.cuh"
#include "optim.cuh"
#include "peer.cuh"
#include "warp_cl.cuh"

namespace {
thread_local char g_err[512] = "";

int fail(mvf_status st, const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess)
        snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else
        snprintf(g_err, sizeof(g_err), "%s", what);
    return (int)st;
}

bool fill_args(mvf::F1Args& a, const mvf_f1_params* p) {
    if (!p || p->B <= 0 || p->H < 3 || p->W < 3) return false;
    a.B = p->B;
    a.H = p->H;
    a.W = p->W;
    a.min_disp = p->min_disp;
    a.disp_range = p->disp_range;
    a.smooth_w = p->smooth_w;
    a.flags = p->flags;
    return true;
}
}  // namespace

extern "C" {

int mvf_version(void) { return 100; }
const char* mvf_last_error(void) { return g_err; }

size_t mvf_f1_workspace_bytes(int B) { return B > 0 ? mvf::f1_workspace_bytes(B) : 0; }

int mvf_workspace_init(void* workspace, size_t bytes, void* stream) {
    if (!workspace || bytes == 0) return fail(MVF_ERR_INVALID, "mvf_workspace_init: null workspace");
    cudaError_t e = cudaMemsetAsync(workspace, 0, bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(MVF_ERR_CUDA, "mvf_workspace_init", e);
    return MVF_OK;
}

int mvf_f1_forward(const mvf_f1_params* p, const float* disp, const float* tgt, const float* src0,
                   const float* src1, const float* inv_K, const float* P0, const float* P1, const float* noise,
                   const float* mask_rec, float* loss, float* stats, uint8_t* idx, int32_t* x0y0, float* warp0,
                   float* warp1, float* to_optimise, void* workspace, size_t workspace_bytes, void* stream) {
    mvf::F1Args a;
    memset(&a, 0, sizeof(a));
    if (!fill_args(a, p)) return fail(MVF_ERR_INVALID, "mvf_f1_forward: bad params (need B>0, H>=3, W>=3)");
    if (!disp || !tgt || !src0 || !src1 || !inv_K || !P0 || !P1 || !loss || !stats || !idx)
        return fail(MVF_ERR_INVALID, "mvf_f1_forward: null required pointer");
    if (x0y0 && (!warp0 || !warp1)) return fail(MVF_ERR_INVALID, "mvf_f1_forward: x0y0 needs warp0/warp1");
    if (!workspace || workspace_bytes < mvf::f1_workspace_bytes(p->B))
        return fail(MVF_ERR_WORKSPACE, "mvf_f1_forward: workspace too small");
    a.disp = disp; a.tgt = tgt; a.src0 = src0; a.src1 = src1; a.inv_K = inv_K; a.P0 = P0; a.P1 = P1;
    a.noise = noise; a.mask = mask_rec; a.loss = loss; a.stats = stats; a.idx = idx;
    a.x0y0 = x0y0; a.warp0 = warp0; a.warp1 = warp1; a.to_opt = to_optimise;
    a.ws = (mvf::F1Workspace*)workspace;
    cudaError_t e = mvf::launch_f1_forward(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(MVF_ERR_CUDA, "mvf_f1_forward launch", e);
    return MVF_OK;
}

int mvf_f1_backward(const mvf_f1_params* p, const float* disp, const float* tgt, const float* src0,
                    const float* src1, const float* inv_K, const float* P0, const float* P1,
                    const float* mask_rec, const uint8_t* idx, const float* stats, const float* gout,
                    float* g_disp, float* g_P0, float* g_P1, void* workspace, size_t workspace_bytes,
                    void* stream) {
    mvf::F1Args a;
    memset(&a, 0, sizeof(a));
    if (!fill_args(a, p)) return fail(MVF_ERR_INVALID, "mvf_f1_backward: bad params (need B>0, H>=3, W>=3)");
    if (!disp || !tgt || !src0 || !src1 || !inv_K || !P0 || !P1 || !idx || !stats || !g_disp || !g_P0 || !g_P1)
        return fail(MVF_ERR_INVALID, "mvf_f1_backward: null required pointer");
    if (!workspace || workspace_bytes < mvf::f1_workspace_bytes(p->B))
        return fail(MVF_ERR_WORKSPACE, "mvf_f1_backward: workspace too small");
    a.disp = disp; a.tgt = tgt; a.src0 = src0; a.src1 = src1; a.inv_K = inv_K; a.P0 = P0; a.P1 = P1;
    a.mask = mask_rec; a.idx = const_cast<uint8_t*>(idx); a.stats = const_cast<float*>(stats); a.gout = gout;
    a.g_disp = g_disp; a.g_P0 = g_P0; a.g_P1 = g_P1;
    a.ws = (mvf::F1Workspace*)workspace;
    cudaError_t e = mvf::launch_f1_backward(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(MVF_ERR_CUDA, "mvf_f1_backward launch", e);
    return MVF_OK;
}

int mvf_f1_forward_host(const mvf_f1_params* p, const float* disp, const float* tgt, const float* src0,
                        const float* src1, const float* inv_K, const float* P0, const float* P1,
                        const float* noise, const float* mask_rec, float* loss, uint8_t* idx) {
    if (!p || p->B <= 0 || p->H < 3 || p->W < 3) return fail(MVF_ERR_INVALID, "mvf_f1_forward_host: bad params");
    if (!disp || !tgt || !src0 || !src1 || !inv_K || !P0 || !P1 || !loss)
        return fail(MVF_ERR_INVALID, "mvf_f1_forward_host: null required pointer");
    const size_t n = (size_t)p->B * p->H * p->W;
    const bool avg = p->flags & MVF_AVG_REPROJECTION, am = !(p->flags & MVF_DISABLE_AUTOMASKING);
    const size_t nid = am ? (avg ? 1 : 2) : 0;
    const size_t fl = n * (1 + 9 + (noise ? nid : 0) + (mask_rec ? 1 : 0)) + (size_t)p->B * (16 + 24 + 4) + 4;
    const size_t wsb = mvf::f1_workspace_bytes(p->B);
    char* base = nullptr;
    cudaError_t e = cudaMalloc(&base, fl * sizeof(float) + n + wsb + 256);
    if (e != cudaSuccess) return fail(MVF_ERR_CUDA, "mvf_f1_forward_host: cudaMalloc", e);
    cudaStream_t st = 0;
    float* f = (float*)base;
    auto put = [&](const float* h, size_t cnt) -> float* {
        float* d = f;
        if (e == cudaSuccess) e = cudaMemcpyAsync(d, h, cnt * sizeof(float), cudaMemcpyHostToDevice, st);
        f += cnt;
        return d;
    };
    float* d_disp = put(disp, n);
    float* d_tgt = put(tgt, 3 * n);
    float* d_s0 = put(src0, 3 * n);
    float* d_s1 = put(src1, 3 * n);
    float* d_noise = (noise && nid) ? put(noise, nid * n) : nullptr;
    float* d_mask = mask_rec ? put(mask_rec, n) : nullptr;
    float* d_ik = put(inv_K, 16 * (size_t)p->B);
    float* d_p0 = put(P0, 12 * (size_t)p->B);
    float* d_p1 = put(P1, 12 * (size_t)p->B);
    float* d_stats = f; f += 4 * (size_t)p->B;
    float* d_loss = f; f += 4;
    uint8_t* d_idx = (uint8_t*)f;
    void* d_ws = (void*)(((uintptr_t)(d_idx + n) + 255) & ~(uintptr_t)255);
    int rc = MVF_OK;
    if (e != cudaSuccess) rc = fail(MVF_ERR_CUDA, "mvf_f1_forward_host: H2D", e);
    if (rc == MVF_OK) rc = mvf_workspace_init(d_ws, wsb, st);
    if (rc == MVF_OK)
        rc = mvf_f1_forward(p, d_disp, d_tgt, d_s0, d_s1, d_ik, d_p0, d_p1, d_noise, d_mask, d_loss, d_stats, d_idx,
                            nullptr, nullptr, nullptr, nullptr, d_ws, wsb, st);
    if (rc == MVF_OK) {
        e = cudaMemcpyAsync(loss, d_loss, 4 * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && idx) e = cudaMemcpyAsync(idx, d_idx, n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(MVF_ERR_CUDA, "mvf_f1_forward_host: D2H", e);
    }
    cudaFree(base);
    return rc;
}

/* ---- unfused layers.py ops ---------------------------------------------------------------------------- */
#define MVF_RUN(name, expr)                                                     \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) return fail(MVF_ERR_CUDA, name, e__);           \
        return MVF_OK;                                                          \
    } while (0)

int mvf_disp_to_depth_fwd(const float* disp, float* scaled_disp, float* depth, size_t n, float min_disp,
                          float disp_range, void* stream) {
    if (!disp || n == 0) return fail(MVF_ERR_INVALID, "mvf_disp_to_depth_fwd: bad argument");
    MVF_RUN("mvf_disp_to_depth_fwd", mvf::disp_to_depth_fwd(disp, scaled_disp, depth, n, min_disp, disp_range, (cudaStream_t)stream));
}
int mvf_disp_to_depth_bwd(const float* disp, const float* g_scaled_disp, const float* g_depth, float* g_disp, size_t n,
                          float min_disp, float disp_range, void* stream) {
    if (!disp || !g_disp || n == 0) return fail(MVF_ERR_INVALID, "mvf_disp_to_depth_bwd: bad argument");
    MVF_RUN("mvf_disp_to_depth_bwd", mvf::disp_to_depth_bwd(disp, g_scaled_disp, g_depth, g_disp, n, min_disp, disp_range, (cudaStream_t)stream));
}
int mvf_backproject_fwd(const float* depth, const float* inv_K, float* cam_points, int B, int H, int W, void* stream) {
    if (!depth || !inv_K || !cam_points || B <= 0 || H <= 0 || W <= 0) return fail(MVF_ERR_INVALID, "mvf_backproject_fwd: bad argument");
    MVF_RUN("mvf_backproject_fwd", mvf::backproject_fwd(depth, inv_K, cam_points, B, H, W, (cudaStream_t)stream));
}
int mvf_backproject_bwd(const float* g_cam_points, const float* inv_K, float* g_depth, int B, int H, int W, void* stream) {
    if (!g_cam_points || !inv_K || !g_depth || B <= 0 || H <= 0 || W <= 0) return fail(MVF_ERR_INVALID, "mvf_backproject_bwd: bad argument");
    MVF_RUN("mvf_backproject_bwd", mvf::backproject_bwd(g_cam_points, inv_K, g_depth, B, H, W, (cudaStream_t)stream));
}
int mvf_project_fwd(const float* points, const float* P, float* pix_coords, int B, int H, int W, float eps, void* stream) {
    if (!points || !P || !pix_coords || B <= 0 || H < 2 || W < 2) return fail(MVF_ERR_INVALID, "mvf_project_fwd: bad argument");
    MVF_RUN("mvf_project_fwd", mvf::project_fwd(points, P, pix_coords, B, H, W, eps, (cudaStream_t)stream));
}
int mvf_project_bwd(const float* points, const float* P, const float* g_pix_coords, float* g_points, float* g_P,
                    void* workspace, size_t workspace_bytes, int B, int H, int W, float eps, void* stream) {
    if (!points || !P || !g_pix_coords || !g_points || !g_P || B <= 0) return fail(MVF_ERR_INVALID, "mvf_project_bwd: bad argument");
    if (!workspace || workspace_bytes < mvf::f1_workspace_bytes(B)) return fail(MVF_ERR_WORKSPACE, "mvf_project_bwd: workspace too small");
    MVF_RUN("mvf_project_bwd", mvf::project_bwd(points, P, g_pix_coords, g_points, g_P, mvf::ws_fwd_acc((mvf::F1Workspace*)workspace), B, H, W, eps, (cudaStream_t)stream));
}
int mvf_ssim_fwd(const float* x, const float* y, float* out, int N, int H, int W, void* stream) {
    if (!x || !y || !out || N <= 0 || H < 2 || W < 2) return fail(MVF_ERR_INVALID, "mvf_ssim_fwd: bad argument");
    MVF_RUN("mvf_ssim_fwd", mvf::ssim_fwd(x, y, out, N, H, W, (cudaStream_t)stream));
}
int mvf_ssim_bwd(const float* x, const float* y, const float* g_out, float* g_x, float* coef_scratch, int N, int H,
                 int W, void* stream) {
    if (!x || !y || !g_out || !g_x || !coef_scratch || N <= 0 || H < 2 || W < 2) return fail(MVF_ERR_INVALID, "mvf_ssim_bwd: bad argument");
    MVF_RUN("mvf_ssim_bwd", mvf::ssim_bwd(x, y, g_out, g_x, coef_scratch, N, H, W, (cudaStream_t)stream));
}
int mvf_smooth_loss_fwd(const float* disp, const float* img, float* loss, void* workspace, size_t workspace_bytes,
                        int B, int H, int W, void* stream) {
    if (!disp || !img || !loss || B <= 0 || H < 2 || W < 2) return fail(MVF_ERR_INVALID, "mvf_smooth_loss_fwd: bad argument");
    if (!workspace || workspace_bytes < mvf::f1_workspace_bytes(B)) return fail(MVF_ERR_WORKSPACE, "mvf_smooth_loss_fwd: workspace too small");
    MVF_RUN("mvf_smooth_loss_fwd", mvf::smooth_fwd(disp, img, loss, mvf::ws_fwd_acc((mvf::F1Workspace*)workspace), B, H, W, (cudaStream_t)stream));
}
int mvf_smooth_loss_bwd(const float* disp, const float* img, const float* gout, float* g_disp, int B, int H, int W,
                        void* stream) {
    if (!disp || !img || !g_disp || B <= 0 || H < 2 || W < 2) return fail(MVF_ERR_INVALID, "mvf_smooth_loss_bwd: bad argument");
    MVF_RUN("mvf_smooth_loss_bwd", mvf::smooth_bwd(disp, img, gout, g_disp, B, H, W, (cudaStream_t)stream));
}
int mvf_si_log_fwd(const float* pred, const float* target, const float* mask, float* loss, float* stats,
                   void* workspace, size_t workspace_bytes, int B, size_t HW, float beta, void* stream) {
    if (!pred || !target || !loss || !stats || B <= 0 || HW == 0) return fail(MVF_ERR_INVALID, "mvf_si_log_fwd: bad argument");
    if (!workspace || workspace_bytes < mvf::f1_workspace_bytes(B)) return fail(MVF_ERR_WORKSPACE, "mvf_si_log_fwd: workspace too small");
    MVF_RUN("mvf_si_log_fwd", mvf::si_log_fwd(pred, target, mask, loss, stats, mvf::ws_fwd_acc((mvf::F1Workspace*)workspace), B, HW, beta, (cudaStream_t)stream));
}
int mvf_si_log_bwd(const float* pred, const float* target, const float* mask, const float* stats, const float* gout,
                   float* g_pred, float* g_target, int B, size_t HW, float beta, void* stream) {
    if (!pred || !target || !stats || B <= 0 || HW == 0) return fail(MVF_ERR_INVALID, "mvf_si_log_bwd: bad argument");
    MVF_RUN("mvf_si_log_bwd", mvf::si_log_bwd(pred, target, mask, stats, gout, g_pred, g_target, B, HW, beta, (cudaStream_t)stream));
}


/* ---- tensor-core convolutions ----------------------------------------------------------------------------- */
size_t mvf_conv2d_packed_filter_floats(int N, int K, int KH, int KW) {
    if (N <= 0 || K <= 0 || KH <= 0 || KW <= 0) return 0;
    return mvf::tc::packed_filter_floats(N, K, KH, KW);
}
int mvf_conv2d_pack_filters(const float* w, float* packed, int Cout, int Cin, int KH, int KW, int dgrad, void* stream) {
    if (!w || !packed || Cout <= 0 || Cin <= 0 || KH <= 0 || KW <= 0) return fail(MVF_ERR_INVALID, "mvf_conv2d_pack_filters: bad argument");
    MVF_RUN("mvf_conv2d_pack_filters", mvf::tc::pack_filters(w, packed, Cout, Cin, KH, KW, dgrad, (cudaStream_t)stream));
}
static mvf::tc::ConvDesc to_desc(const mvf_conv2d_desc* d) {
    mvf::tc::ConvDesc c;
    c.B = d->B; c.Cin = d->Cin; c.H = d->H; c.W = d->W; c.Cout = d->Cout; c.KH = d->KH; c.KW = d->KW; c.pad = d->pad;
    c.stride = d->stride;
    c.stride_x = d->stride_x > 0 ? d->stride_x : d->stride;
    c.x_sB = d->x_stride[0]; c.x_sH = d->x_stride[1]; c.x_sW = d->x_stride[2];
    c.y_sB = d->y_stride[0]; c.y_sH = d->y_stride[1]; c.y_sW = d->y_stride[2];
    return c;
}
int mvf_selftest_umma(const float* A, const float* B, float* D, int N, int K, int a_mn_major, void* stream) {
    if (!A || !B || !D || N < 16 || N > 256 || N % 16 || K < 32 || K % 32) return fail(MVF_ERR_INVALID, "mvf_selftest_umma: bad argument");
    if (a_mn_major) return fail(MVF_ERR_INVALID, "mvf_selftest_umma: use mvf_selftest_umma_rows(mode 2) for the MN-major operand");
    MVF_RUN("mvf_selftest_umma", mvf::tc::umma_selftest(A, B, D, N, K, a_mn_major, (cudaStream_t)stream));
}
int mvf_selftest_umma_rows(const float* A, const float* B, float* D, int N, int K, int row_off, int base_off_mode, void* stream) {
    // base_off_mode: 0 / 1 = K-major A with rows shifted (descriptor base-offset field clear / set);
    //                2 = MN-major A ([K+8][128]) with the k-rows shifted by row_off (<= 8)
    if (!A || !B || !D || N < 16 || N > 256 || N % 16 || K < 32 || K % 32 || row_off < 0 || row_off > 32 || (base_off_mode == 2 && row_off > 8))
        return fail(MVF_ERR_INVALID, "mvf_selftest_umma_rows: bad argument");
    MVF_RUN("mvf_selftest_umma_rows", mvf::tc::umma_selftest(A, B, D, N, K, base_off_mode == 2 ? 1 : 0, (cudaStream_t)stream, row_off,
                                                             base_off_mode == 2 ? 0 : base_off_mode));
}
void mvf_conv2d_debug_buffer(float* p) { mvf::tc::set_debug_buffer(p); }
int mvf_conv2d_supported(const mvf_conv2d_desc* d) {
    if (!d) return 0;
    const char* why = mvf::tc::conv_check(to_desc(d));
    if (why) {
        fail(MVF_ERR_INVALID, why);
        return 0;
    }
    return 1;
}
int mvf_conv2d_forward(const mvf_conv2d_desc* d, const float* x, const float* w_packed, const float* bias, float* y,
                       int act, void* stream) {
    if (!d || !x || !w_packed || !y) return fail(MVF_ERR_INVALID, "mvf_conv2d_forward: null pointer");
    const mvf::tc::ConvDesc c = to_desc(d);
    const char* why = mvf::tc::conv_check(c);
    if (why) return fail(MVF_ERR_INVALID, why);
    cudaError_t e = mvf::tc::conv_forward(c, x, w_packed, bias, y, act, (cudaStream_t)stream, &why);
    if (e != cudaSuccess) return why ? fail(MVF_ERR_CUDA, why) : fail(MVF_ERR_CUDA, "mvf_conv2d_forward launch", e);
    return MVF_OK;
}
int mvf_conv2d_pack_chunk(void) { return mvf::tc::pack_chunk(); }
int mvf_conv2d_pack_filters_multi(const long long* table, int n_entries, long long total_blocks, void* stream) {
    if (!table || n_entries <= 0 || total_blocks <= 0 || total_blocks > 0x7fffffffLL) return fail(MVF_ERR_INVALID, "mvf_conv2d_pack_filters_multi: bad argument");
    MVF_RUN("mvf_conv2d_pack_filters_multi", mvf::tc::pack_filters_multi(table, n_entries, total_blocks, (cudaStream_t)stream));
}
int mvf_conv2d_forward_prelu(const mvf_conv2d_desc* d, const float* x, const float* w_packed, const float* bias, const float* slope,
                             float* y, void* stream) {
    if (!d || !x || !w_packed || !y || !slope) return fail(MVF_ERR_INVALID, "mvf_conv2d_forward_prelu: null pointer");
    const mvf::tc::ConvDesc c = to_desc(d);
    const char* why = mvf::tc::conv_check(c);
    if (why) return fail(MVF_ERR_INVALID, why);
    cudaError_t e = mvf::tc::conv_forward(c, x, w_packed, bias, y, 3, (cudaStream_t)stream, &why, slope);
    if (e != cudaSuccess) return why ? fail(MVF_ERR_CUDA, why) : fail(MVF_ERR_CUDA, "mvf_conv2d_forward_prelu launch", e);
    return MVF_OK;
}
int mvf_conv_transpose2d_s2_fwd(const mvf_conv2d_desc* d, const float* x, const float* w_packed, const float* bias, float* y, void* stream) {
    if (!d || !x || !w_packed || !y) return fail(MVF_ERR_INVALID, "mvf_conv_transpose2d_s2_fwd: null pointer");
    const char* why = nullptr;
    cudaError_t e = mvf::tc::conv_dgrad_s2(to_desc(d), x, w_packed, y, (cudaStream_t)stream, &why, bias);
    if (e != cudaSuccess)
        return why ? fail(e == cudaErrorInvalidValue ? MVF_ERR_INVALID : MVF_ERR_CUDA, why) : fail(MVF_ERR_CUDA, "mvf_conv_transpose2d_s2_fwd launch", e);
    return MVF_OK;
}
int mvf_conv2d_dgrad_s2_supported(const mvf_conv2d_desc* d) {
    if (!d) return 0;
    const char* why = mvf::tc::conv_dgrad_s2_check(to_desc(d));
    if (why) {
        fail(MVF_ERR_INVALID, why);
        return 0;
    }
    return 1;
}
int mvf_conv2d_dgrad_s2(const mvf_conv2d_desc* d, const float* grad_y, const float* w_packed, float* grad_x, void* stream) {
    if (!d || !grad_y || !w_packed || !grad_x) return fail(MVF_ERR_INVALID, "mvf_conv2d_dgrad_s2: null pointer");
    const char* why = nullptr;
    cudaError_t e = mvf::tc::conv_dgrad_s2(to_desc(d), grad_y, w_packed, grad_x, (cudaStream_t)stream, &why);
    if (e != cudaSuccess) return why ? fail(e == cudaErrorInvalidValue ? MVF_ERR_INVALID : MVF_ERR_CUDA, why) : fail(MVF_ERR_CUDA, "mvf_conv2d_dgrad_s2 launch", e);
    return MVF_OK;
}
int mvf_conv2d_dgrad_s2_plan(int KH, int KW, int pad, int* table, int capacity) {
    if (!table || capacity <= 0 || KH <= 0 || KW <= 0 || KH > 8 || KW > 8 || pad < 0) return -1;
    return mvf::tc::conv_dgrad_s2_plan_table(KH, KW, pad, table, capacity);
}
static mvf::tc::WgradDesc to_wdesc(const mvf_conv2d_desc* d) {
    mvf::tc::WgradDesc c;
    c.B = d->B; c.Cin = d->Cin; c.H = d->H; c.W = d->W; c.Cout = d->Cout; c.KH = d->KH; c.KW = d->KW; c.pad = d->pad;
    c.stride = d->stride;
    c.stride_x = d->stride_x > 0 ? d->stride_x : d->stride;
    c.x_sB = d->x_stride[0]; c.x_sH = d->x_stride[1]; c.x_sW = d->x_stride[2];
    c.g_sB = d->y_stride[0]; c.g_sH = d->y_stride[1]; c.g_sW = d->y_stride[2];
    return c;
}
int mvf_conv2d_wgrad_supported(const mvf_conv2d_desc* d) {
    if (!d) return 0;
    const char* why = mvf::tc::wgrad_check(to_wdesc(d));
    if (why) {
        fail(MVF_ERR_INVALID, why);
        return 0;
    }
    return 1;
}
size_t mvf_conv2d_wgrad_workspace_floats(const mvf_conv2d_desc* d) {
    if (!d || mvf::tc::wgrad_check(to_wdesc(d))) return 0;
    return mvf::tc::wgrad_workspace_floats(to_wdesc(d));
}
int mvf_conv2d_wgrad(const mvf_conv2d_desc* d, const float* x, const float* grad_out, float* grad_w, float* workspace,
                     size_t workspace_floats, void* stream) {
    if (!d || !x || !grad_out || !grad_w || !workspace) return fail(MVF_ERR_INVALID, "mvf_conv2d_wgrad: null pointer");
    const mvf::tc::WgradDesc c = to_wdesc(d);
    const char* why = mvf::tc::wgrad_check(c);
    if (why) return fail(MVF_ERR_INVALID, why);
    if (workspace_floats < mvf::tc::wgrad_workspace_floats(c)) return fail(MVF_ERR_WORKSPACE, "mvf_conv2d_wgrad: workspace too small");
    cudaError_t e = mvf::tc::conv_wgrad(c, x, grad_out, grad_w, workspace, (cudaStream_t)stream, &why);
    if (e != cudaSuccess) return why ? fail(MVF_ERR_CUDA, why) : fail(MVF_ERR_CUDA, "mvf_conv2d_wgrad launch", e);
    return MVF_OK;
}

/* ---- fused upsample + concat + reflection pad (channels-last) --------------------------------------------------- */
int mvf_upcat_pad_fwd(const float* a, const float* skip, float* y, int B, int Ca, int Cs, int H, int W, int upsample, void* stream) {
    if (!a || !y || B <= 0 || Ca <= 0 || Cs < 0 || (Cs > 0 && !skip) || H < 4 || W < 4 || (Ca % 4) || (Cs % 4) ||
        (upsample && ((H | W) & 1)))
        return fail(MVF_ERR_INVALID, "mvf_upcat_pad_fwd: need channels % 4 == 0, H, W >= 4 (even when upsampling)");
    MVF_RUN("mvf_upcat_pad_fwd", mvf::upcat_pad_fwd(a, skip, y, B, Ca, Cs, H, W, upsample, (cudaStream_t)stream));
}
int mvf_upcat_pad_bwd(const float* grad_y, float* grad_a, float* grad_skip, int B, int Ca, int Cs, int H, int W, int upsample,
                      void* stream) {
    if (!grad_y || B <= 0 || Ca <= 0 || Cs < 0 || H < 4 || W < 4 || (Ca % 4) || (Cs % 4) || (upsample && ((H | W) & 1)))
        return fail(MVF_ERR_INVALID, "mvf_upcat_pad_bwd: need channels % 4 == 0, H, W >= 4 (even when upsampling)");
    MVF_RUN("mvf_upcat_pad_bwd", mvf::upcat_pad_bwd(grad_y, grad_a, grad_skip, B, Ca, Cs, H, W, upsample, (cudaStream_t)stream));
}

/* ---- MaxPool2d(3, 2, 1), channels-last ------------------------------------------------------------------------ */
int mvf_maxpool3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int C, int H, int W, void* stream) {
    if (!x || !y || !idx || B <= 0 || C <= 0 || (C % 4) || H < 2 || W < 2) return fail(MVF_ERR_INVALID, "mvf_maxpool3s2_fwd: bad argument (C % 4 == 0)");
    MVF_RUN("mvf_maxpool3s2_fwd", mvf::maxpool3s2_fwd(x, y, idx, B, C, H, W, (cudaStream_t)stream));
}
int mvf_maxpool3s2_bwd(const float* grad_y, const unsigned char* idx, float* grad_x, int B, int C, int H, int W, void* stream) {
    if (!grad_y || !idx || !grad_x || B <= 0 || C <= 0 || (C % 4) || H < 2 || W < 2) return fail(MVF_ERR_INVALID, "mvf_maxpool3s2_bwd: bad argument (C % 4 == 0)");
    MVF_RUN("mvf_maxpool3s2_bwd", mvf::maxpool3s2_bwd(grad_y, idx, grad_x, B, C, H, W, (cudaStream_t)stream));
}

/* ---- fused BatchNorm2d (training) + residual add + ReLU, channels-last ------------------------------------------- */
static bool bn_shape_ok(long long P, int C) { return P > 0 && C > 0 && C <= 1024 && (C % 4 == 0 || (C % 2 == 0 && P % 2 == 0)); }
size_t mvf_bn_workspace_floats(long long P, int C) { return bn_shape_ok(P, C) ? mvf::bn_workspace_floats(P, C) : 0; }
int mvf_bn_relu_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, float* save_mean, float* save_invstd, float* workspace,
                    size_t workspace_floats, long long P, int C, float eps, float momentum, int relu, void* stream) {
    if (!x || !y || !gamma || !beta || !save_mean || !save_invstd || !workspace || !bn_shape_ok(P, C))
        return fail(MVF_ERR_INVALID, "mvf_bn_relu_fwd: bad argument (C % 4 == 0, or C % 2 == 0 with an even pixel count; C <= 1024)");
    if (workspace_floats < mvf::bn_workspace_floats(P, C)) return fail(MVF_ERR_WORKSPACE, "mvf_bn_relu_fwd: workspace too small");
    MVF_RUN("mvf_bn_relu_fwd", mvf::bn_forward(x, identity, y, gamma, beta, running_mean, running_var, num_batches_tracked, save_mean, save_invstd, workspace,
                                               P, C, eps, momentum, relu, (cudaStream_t)stream));
}
int mvf_bn_relu_bwd(const float* x, const float* grad_y, const float* y, const float* gamma, const float* save_mean,
                    const float* save_invstd, float* grad_x, float* grad_identity, float* grad_gamma, float* grad_beta,
                    float* workspace, size_t workspace_floats, long long P, int C, int relu, void* stream) {
    if (!x || !grad_y || !gamma || !save_mean || !save_invstd || !grad_x || !grad_gamma || !grad_beta || !workspace || !bn_shape_ok(P, C) ||
        (relu && !y))
        return fail(MVF_ERR_INVALID, "mvf_bn_relu_bwd: bad argument (C % 4 == 0, or C % 2 == 0 with an even pixel count; C <= 1024)");
    if (workspace_floats < mvf::bn_workspace_floats(P, C)) return fail(MVF_ERR_WORKSPACE, "mvf_bn_relu_bwd: workspace too small");
    MVF_RUN("mvf_bn_relu_bwd", mvf::bn_backward(x, grad_y, y, gamma, save_mean, save_invstd, grad_x, grad_identity, grad_gamma, grad_beta,
                                                workspace, P, C, relu, (cudaStream_t)stream));
}

/* ---- SyncBatchNorm: the same arithmetic split around a cross-rank sum of [2C + 1] doubles --------------------------- */
int mvf_bn_sync_stats_fwd(const float* x, double* sums, float* workspace, size_t workspace_floats, long long P, int C, void* stream) {
    if (!x || !sums || !workspace || P <= 0 || C <= 0 || (C % 4) || C > 1024)
        return fail(MVF_ERR_INVALID, "mvf_bn_sync_stats_fwd: bad argument (C % 4 == 0, C <= 1024)");
    if (workspace_floats < mvf::bn_workspace_floats(P, C)) return fail(MVF_ERR_WORKSPACE, "mvf_bn_sync_stats_fwd: workspace too small");
    MVF_RUN("mvf_bn_sync_stats_fwd", mvf::bn_sync_stats_fwd(x, sums, workspace, P, C, (cudaStream_t)stream));
}
int mvf_bn_sync_apply_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long* num_batches_tracked, float* save_mean, float* save_invstd,
                          const double* global_sums, long long P, int C, float eps, float momentum, int relu, void* stream) {
    if (!x || !y || !gamma || !beta || !save_mean || !save_invstd || !global_sums || P <= 0 || C <= 0 || (C % 4) || C > 1024)
        return fail(MVF_ERR_INVALID, "mvf_bn_sync_apply_fwd: bad argument (C % 4 == 0, C <= 1024)");
    MVF_RUN("mvf_bn_sync_apply_fwd", mvf::bn_sync_apply_fwd(x, identity, y, gamma, beta, running_mean, running_var, num_batches_tracked,
                                                            save_mean, save_invstd, global_sums, P, C, eps, momentum, relu,
                                                            (cudaStream_t)stream));
}
int mvf_bn_sync_stats_bwd(const float* x, const float* grad_y, const float* y, const float* save_mean, const float* save_invstd,
                          double* sums, float* grad_gamma, float* grad_beta, float* workspace, size_t workspace_floats, long long P,
                          int C, int relu, void* stream) {
    if (!x || !grad_y || !save_mean || !save_invstd || !sums || !grad_gamma || !grad_beta || !workspace || P <= 0 || C <= 0 ||
        (C % 4) || C > 1024 || (relu && !y))
        return fail(MVF_ERR_INVALID, "mvf_bn_sync_stats_bwd: bad argument (C % 4 == 0, C <= 1024)");
    if (workspace_floats < mvf::bn_workspace_floats(P, C)) return fail(MVF_ERR_WORKSPACE, "mvf_bn_sync_stats_bwd: workspace too small");
    MVF_RUN("mvf_bn_sync_stats_bwd", mvf::bn_sync_stats_bwd(x, grad_y, y, save_mean, save_invstd, sums, grad_gamma, grad_beta, workspace,
                                                            P, C, relu, (cudaStream_t)stream));
}
int mvf_bn_sync_apply_bwd(const float* x, const float* grad_y, const float* y, const float* gamma, const float* save_mean,
                          const float* save_invstd, float* grad_x, float* grad_identity, const double* global_sums, float* scratch,
                          size_t scratch_floats, long long P, int C, int relu, void* stream) {
    if (!x || !grad_y || !gamma || !save_mean || !save_invstd || !grad_x || !global_sums || !scratch || P <= 0 || C <= 0 || (C % 4) ||
        C > 1024 || (relu && !y))
        return fail(MVF_ERR_INVALID, "mvf_bn_sync_apply_bwd: bad argument (C % 4 == 0, C <= 1024)");
    if (scratch_floats < (size_t)(2 * C + 4)) return fail(MVF_ERR_WORKSPACE, "mvf_bn_sync_apply_bwd: scratch needs 2C + 4 floats");
    MVF_RUN("mvf_bn_sync_apply_bwd", mvf::bn_sync_apply_bwd(x, grad_y, y, gamma, save_mean, save_invstd, grad_x, grad_identity, global_sums,
                                                            scratch, P, C, relu, (cudaStream_t)stream));
}

int mvf_act_bwd_bias(const float* grad_y, const float* y, float* grad_pre, float* grad_bias, float* workspace, size_t workspace_floats,
                     long long P, int C, int act, void* stream) {
    if (!grad_y || P <= 0 || C <= 0 || (C % 4) || C > 1024 || act < 0 || act > 2 || (act && (!y || !grad_pre)) || (!grad_pre && !grad_bias))
        return fail(MVF_ERR_INVALID, "mvf_act_bwd_bias: bad argument (C % 4 == 0, C <= 1024, act in 0..2)");
    if (grad_bias && (!workspace || workspace_floats < mvf::bn_workspace_floats(P, C)))
        return fail(MVF_ERR_WORKSPACE, "mvf_act_bwd_bias: workspace too small");
    MVF_RUN("mvf_act_bwd_bias", mvf::act_bwd_bias(grad_y, y, grad_pre, grad_bias, workspace, P, C, act, (cudaStream_t)stream));
}

unsigned long long mvf_stream_capture_id(void* stream) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo((cudaStream_t)stream, &status, &id) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return status == cudaStreamCaptureStatusActive ? id : 0;
}

/* ---- fused clip_grad_norm_ + AdamW over a flat arena -------------------------------------------------------------- */
size_t mvf_adamw_workspace_bytes(void) { return mvf::adamw_workspace_bytes(); }
int mvf_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, long long n_duplicated,
                   const unsigned char* skip_groups, float* state, void* workspace, size_t workspace_bytes, float beta1, float beta2,
                   float eps, float weight_decay, float max_norm, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !state || !workspace || n <= 0) return fail(MVF_ERR_INVALID, "mvf_adamw_step: bad argument");
    if (n_duplicated < 0 || n_duplicated > n || (n_duplicated & 3) != 0)
        return fail(MVF_ERR_INVALID, "mvf_adamw_step: n_duplicated must be a multiple of 4 in [0, n]");
    if (workspace_bytes < mvf::adamw_workspace_bytes()) return fail(MVF_ERR_WORKSPACE, "mvf_adamw_step: workspace too small");
    if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
        return fail(MVF_ERR_INVALID, "mvf_adamw_step: arenas must be 16-byte aligned");
    MVF_RUN("mvf_adamw_step", mvf::adamw_step(params, grads, exp_avg, exp_avg_sq, n, n_duplicated, skip_groups, state, workspace,
                                              beta1, beta2, eps, weight_decay, max_norm, (cudaStream_t)stream));
}

int mvf_gather_grads(float* arena, const void* const* grads, const long long* offsets, const long long* sizes, int n_tensors,
                     void* stream) {
    if (!arena || !grads || !offsets || !sizes || n_tensors <= 0) return fail(MVF_ERR_INVALID, "mvf_gather_grads: bad argument");
    if (((uintptr_t)arena & 15) != 0) return fail(MVF_ERR_INVALID, "mvf_gather_grads: arena must be 16-byte aligned");
    for (int i = 0; i < n_tensors; ++i)
        if (offsets[i] < 0 || (offsets[i] & 3) != 0 || sizes[i] < 0) return fail(MVF_ERR_INVALID, "mvf_gather_grads: offsets must be non-negative multiples of 4");
    MVF_RUN("mvf_gather_grads", mvf::gather_grads(arena, grads, offsets, sizes, n_tensors, (cudaStream_t)stream));
}

int mvf_flow_warp_fwd(const float* x, const float* flow, float* y, int B, int C, int H, int W, int layout, void* stream) {
    if (!x || !flow || !y || B <= 0 || C <= 0 || H < 2 || W < 2 || (layout != 0 && layout != 1) || (layout == 1 && (C % 2)))
        return fail(MVF_ERR_INVALID, "mvf_flow_warp_fwd: bad argument (H, W >= 2; channels-last needs C % 2 == 0)");
    MVF_RUN("mvf_flow_warp_fwd", mvf::flow_warp_fwd(x, flow, y, B, C, H, W, layout, (cudaStream_t)stream));
}
size_t mvf_flow_warp_bwd_workspace_bytes(int B, int C, int H, int W) {
    return (B > 0 && C > 0 && H > 0 && W > 0) ? mvf::flow_warp_bwd_workspace_bytes(B, C, H, W) : 0;
}
int mvf_flow_warp_bwd(const float* grad_y, const float* flow, float* grad_x, int B, int C, int H, int W, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (!grad_y || !flow || !grad_x || !workspace || B <= 0 || C <= 0 || (C % 2) || H < 2 || W < 2)
        return fail(MVF_ERR_INVALID, "mvf_flow_warp_bwd: bad argument (channels-last, C % 2 == 0, H, W >= 2)");
    if (workspace_bytes < mvf::flow_warp_bwd_workspace_bytes(B, C, H, W) || ((uintptr_t)workspace & 15))
        return fail(MVF_ERR_INVALID, "mvf_flow_warp_bwd: workspace too small or not 16-byte aligned");
    MVF_RUN("mvf_flow_warp_bwd", mvf::flow_warp_bwd(grad_y, flow, grad_x, B, C, H, W, workspace, workspace_bytes, (cudaStream_t)stream));
}
static bool resize_args_ok(const void* a, const void* b, int B, int C, int Hi, int Wi, int Ho, int Wo, int layout) {
    return a && b && B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && (layout == 0 || (layout == 1 && C % 2 == 0));
}
int mvf_resize_bilinear_fwd(const float* x, float* y, int B, int C, int Hin, int Win, int Hout, int Wout, float scale_h, float scale_w,
                            int align_corners, float mul_even, float mul_odd, int layout, void* stream) {
    if (!resize_args_ok(x, y, B, C, Hin, Win, Hout, Wout, layout)) return fail(MVF_ERR_INVALID, "mvf_resize_bilinear_fwd: bad argument");
    if (layout == 1 && (mul_even != 1.f || mul_odd != 1.f)) return fail(MVF_ERR_INVALID, "mvf_resize_bilinear_fwd: multipliers are NCHW-only");
    MVF_RUN("mvf_resize_bilinear_fwd", mvf::resize_bilinear_fwd(x, y, B, C, Hin, Win, Hout, Wout, scale_h, scale_w, align_corners, mul_even,
                                                                mul_odd, layout, (cudaStream_t)stream));
}
int mvf_resize_bilinear_bwd(const float* grad_y, float* grad_x, int B, int C, int Hin, int Win, int Hout, int Wout, float scale_h,
                            float scale_w, int align_corners, float mul_even, float mul_odd, int layout, void* stream) {
    if (!resize_args_ok(grad_y, grad_x, B, C, Hin, Win, Hout, Wout, layout)) return fail(MVF_ERR_INVALID, "mvf_resize_bilinear_bwd: bad argument");
    if (layout == 1 && (mul_even != 1.f || mul_odd != 1.f)) return fail(MVF_ERR_INVALID, "mvf_resize_bilinear_bwd: multipliers are NCHW-only");
    MVF_RUN("mvf_resize_bilinear_bwd", mvf::resize_bilinear_bwd(grad_y, grad_x, B, C, Hin, Win, Hout, Wout, scale_h, scale_w, align_corners,
                                                                mul_even, mul_odd, layout, (cudaStream_t)stream));
}
int mvf_prelu_cl_fwd(const float* x, const float* res, const float* slope, float* y, long long P, int C, void* stream) {
    if (!x || !slope || !y || P <= 0 || C <= 0 || (C % 4) || ((uintptr_t)slope & 15))
        return fail(MVF_ERR_INVALID, "mvf_prelu_cl_fwd: bad argument (C % 4 == 0, 16-byte aligned slope)");
    MVF_RUN("mvf_prelu_cl_fwd", mvf::prelu_cl_fwd(x, res, slope, y, P, C, (cudaStream_t)stream));
}
int mvf_pose_matrix_fwd(const float* axisangle, const float* translation, float* M, int B, int invert, void* stream) {
    if (!axisangle || !translation || !M || B <= 0) return fail(MVF_ERR_INVALID, "mvf_pose_matrix_fwd: bad argument");
    MVF_RUN("mvf_pose_matrix_fwd", mvf::pose_matrix_fwd(axisangle, translation, M, B, invert, (cudaStream_t)stream));
}
int mvf_pose_matrix_bwd(const float* axisangle, const float* translation, const float* grad_M, float* grad_axisangle,
                        float* grad_translation, int B, int invert, void* stream) {
    if (!axisangle || !translation || !grad_M || !grad_axisangle || !grad_translation || B <= 0)
        return fail(MVF_ERR_INVALID, "mvf_pose_matrix_bwd: bad argument");
    MVF_RUN("mvf_pose_matrix_bwd", mvf::pose_matrix_bwd(axisangle, translation, grad_M, grad_axisangle, grad_translation, B, invert,
                                                        (cudaStream_t)stream));
}

int mvf_dwconv3x3_fwd(const float* x, const float* w_taps, const float* bias, float* y, int B, int C, int H, int W, int dilation, int flip,
                      void* stream) {
    if (!x || !w_taps || !y || B <= 0 || C <= 0 || (C % 4) || H <= 0 || W <= 0 || dilation < 1)
        return fail(MVF_ERR_INVALID, "mvf_dwconv3x3_fwd: bad argument (C % 4 == 0, dilation >= 1)");
    MVF_RUN("mvf_dwconv3x3_fwd", mvf::dwconv3x3_fwd(x, w_taps, bias, y, B, C, H, W, dilation, flip, (cudaStream_t)stream));
}
size_t mvf_dwconv3x3_wgrad_workspace_floats(long long P, int C) { return (P > 0 && C > 0) ? mvf::dwconv3x3_wgrad_workspace_floats(P, C) : 0; }
int mvf_dwconv3x3_wgrad(const float* x, const float* grad_y, float* grad_w, float* workspace, size_t workspace_floats, int B, int C, int H,
                        int W, int dilation, void* stream) {
    if (!x || !grad_y || !grad_w || !workspace || B <= 0 || C <= 0 || (C % 4) || C > 1024 || H <= 0 || W <= 0 || dilation < 1)
        return fail(MVF_ERR_INVALID, "mvf_dwconv3x3_wgrad: bad argument (C % 4 == 0, C <= 1024)");
    if (workspace_floats < mvf::dwconv3x3_wgrad_workspace_floats((long long)B * H * W, C)) return fail(MVF_ERR_INVALID, "mvf_dwconv3x3_wgrad: workspace too small");
    MVF_RUN("mvf_dwconv3x3_wgrad", mvf::dwconv3x3_wgrad(x, grad_y, grad_w, workspace, B, C, H, W, dilation, (cudaStream_t)stream));
}
int mvf_gelu_fwd(const float* x, float* y, long long n, void* stream) {
    if (!x || !y || n <= 0 || (n % 4)) return fail(MVF_ERR_INVALID, "mvf_gelu_fwd: bad argument (n % 4 == 0)");
    MVF_RUN("mvf_gelu_fwd", mvf::gelu_fwd(x, y, n, (cudaStream_t)stream));
}
int mvf_gelu_bwd(const float* x, const float* grad_y, float* grad_x, long long n, void* stream) {
    if (!x || !grad_y || !grad_x || n <= 0 || (n % 4)) return fail(MVF_ERR_INVALID, "mvf_gelu_bwd: bad argument (n % 4 == 0)");
    MVF_RUN("mvf_gelu_bwd", mvf::gelu_bwd(x, grad_y, grad_x, n, (cudaStream_t)stream));
}
int mvf_layernorm_cl_fwd(const float* x, const float* weight, const float* bias, float* y, float* mean, float* rstd, long long P, int C,
                         float eps, void* stream) {
    if (!x || !weight || !bias || !y || !mean || !rstd || P <= 0 || C <= 0 || (C % 4) || C > 512)
        return fail(MVF_ERR_INVALID, "mvf_layernorm_cl_fwd: bad argument (C % 4 == 0, C <= 512)");
    MVF_RUN("mvf_layernorm_cl_fwd", mvf::layernorm_cl_fwd(x, weight, bias, y, mean, rstd, P, C, eps, (cudaStream_t)stream));
}
size_t mvf_layernorm_bwd_workspace_floats(long long P, int C) { return (P > 0 && C > 0) ? mvf::layernorm_bwd_workspace_floats(P, C) : 0; }
int mvf_layernorm_cl_bwd(const float* x, const float* grad_y, const float* weight, const float* mean, const float* rstd, float* grad_x,
                         float* grad_weight, float* grad_bias, float* workspace, size_t workspace_floats, long long P, int C, void* stream) {
    if (!x || !grad_y || !weight || !mean || !rstd || !grad_x || !grad_weight || !grad_bias || !workspace || P <= 0 || C <= 0 || (C % 4) || C > 512)
        return fail(MVF_ERR_INVALID, "mvf_layernorm_cl_bwd: bad argument (C % 4 == 0, C <= 512)");
    if (workspace_floats < mvf::layernorm_bwd_workspace_floats(P, C)) return fail(MVF_ERR_INVALID, "mvf_layernorm_cl_bwd: workspace too small");
    MVF_RUN("mvf_layernorm_cl_bwd", mvf::layernorm_cl_bwd(x, grad_y, weight, mean, rstd, grad_x, grad_weight, grad_bias, workspace, P, C,
                                                           (cudaStream_t)stream));
}

size_t mvf_peer_buffer_bytes(void) { return mvf::peer_buffer_bytes(); }
int mvf_peer_allreduce_f64(double* vec, int n, void* const* peers_dev, int rank, int world, int channel, unsigned long long* seq_local,
                           void* stream) {
    if (!vec || !peers_dev || !seq_local || n <= 0 || n > mvf::PEER_MAX_N || world < 1 || world > mvf::PEER_MAX_WORLD || rank < 0 ||
        rank >= world || channel < 0 || channel >= mvf::PEER_CHANNELS)
        return fail(MVF_ERR_INVALID, "mvf_peer_allreduce_f64: bad argument (n <= 2056, world <= 16, channel < 8)");
    MVF_RUN("mvf_peer_allreduce_f64", mvf::peer_allreduce_f64(vec, n, peers_dev, rank, world, channel, seq_local, (cudaStream_t)stream));
}

int mvf_bn_eval_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, long long P, int C, float eps, int relu, void* stream) {
    if (!x || !y || !gamma || !beta || !running_mean || !running_var || P <= 0 || C <= 0 || (C % 4))
        return fail(MVF_ERR_INVALID, "mvf_bn_eval_fwd: bad argument (C % 4 == 0)");
    MVF_RUN("mvf_bn_eval_fwd", mvf::bn_eval_fwd(x, identity, y, gamma, beta, running_mean, running_var, P, C, eps, relu, (cudaStream_t)stream));
}
size_t mvf_depth_eval_workspace_bytes(int Hg, int Wg) { return (Hg > 0 && Wg > 0) ? mvf::depth_eval_workspace_bytes(Hg, Wg) : 0; }
int mvf_depth_eval(const float* disp, int h, int w, const float* gt, int Hg, int Wg, float min_depth, float max_depth, int eigen_crop,
                   float stereo_scale, void* workspace, size_t workspace_bytes, float* metrics8, void* stream) {
    if (!disp || !gt || !workspace || !metrics8 || h <= 0 || w <= 0 || Hg <= 0 || Wg <= 0 || ((uintptr_t)workspace & 15))
        return fail(MVF_ERR_INVALID, "mvf_depth_eval: bad argument");
    if (workspace_bytes < mvf::depth_eval_workspace_bytes(Hg, Wg)) return fail(MVF_ERR_WORKSPACE, "mvf_depth_eval: workspace too small");
    MVF_RUN("mvf_depth_eval", mvf::depth_eval(disp, h, w, gt, Hg, Wg, min_depth, max_depth, eigen_crop, stereo_scale, workspace, metrics8,
                                              (cudaStream_t)stream));
}

size_t mvf_input_pipeline_workspace_floats(int B, int F) { return (B > 0 && F > 0) ? mvf::input_pipeline_workspace_floats(B, F) : 0; }
int mvf_input_pipeline(const unsigned char* frames, const float* prm_f, const int* prm_i, float* workspace, size_t workspace_floats,
                       float* const* color_dev, float* const* color_aug_dev, int B, int F, int H, int W, void* stream) {
    if (!frames || !prm_f || !prm_i || !workspace || !color_dev || !color_aug_dev || B <= 0 || F <= 0 || F > 16 || H <= 0 || W <= 0)
        return fail(MVF_ERR_INVALID, "mvf_input_pipeline: bad argument");
    if (workspace_floats < mvf::input_pipeline_workspace_floats(B, F)) return fail(MVF_ERR_WORKSPACE, "mvf_input_pipeline: workspace too small");
    MVF_RUN("mvf_input_pipeline", mvf::input_pipeline(frames, prm_f, prm_i, workspace, color_dev, color_aug_dev, B, F, H, W, (cudaStream_t)stream));
}

static bool dispconv_ok(int B, int C, int H, int W) { return B > 0 && C > 0 && C % 4 == 0 && C <= 64 && H > 0 && W > 0; }
int mvf_dispconv_fwd(const float* xp, const float* w, const float* bias, float* y, int B, int C, int H, int W, void* stream) {
    if (!xp || !w || !y || !dispconv_ok(B, C, H, W)) return fail(MVF_ERR_INVALID, "mvf_dispconv_fwd: bad argument (C % 4 == 0, C <= 64)");
    MVF_RUN("mvf_dispconv_fwd", mvf::dispconv_fwd(xp, w, bias, y, B, C, H, W, (cudaStream_t)stream));
}
int mvf_dispconv_dgrad(const float* grad_y, const float* w, float* grad_xp, int B, int C, int H, int W, void* stream) {
    if (!grad_y || !w || !grad_xp || !dispconv_ok(B, C, H, W)) return fail(MVF_ERR_INVALID, "mvf_dispconv_dgrad: bad argument (C % 4 == 0, C <= 64)");
    MVF_RUN("mvf_dispconv_dgrad", mvf::dispconv_dgrad(grad_y, w, grad_xp, B, C, H, W, (cudaStream_t)stream));
}
size_t mvf_dispconv_wgrad_workspace_floats(long long P, int C) { return (P > 0 && C > 0) ? mvf::dispconv_wgrad_workspace_floats(P, C) : 0; }
int mvf_dispconv_wgrad(const float* xp, const float* grad_y, float* grad_w, float* grad_b, float* workspace, size_t workspace_floats, int B,
                       int C, int H, int W, void* stream) {
    if (!xp || !grad_y || !grad_w || !workspace || !dispconv_ok(B, C, H, W)) return fail(MVF_ERR_INVALID, "mvf_dispconv_wgrad: bad argument (C % 4 == 0, C <= 64)");
    if (workspace_floats < mvf::dispconv_wgrad_workspace_floats((long long)B * H * W, C)) return fail(MVF_ERR_WORKSPACE, "mvf_dispconv_wgrad: workspace too small");
    MVF_RUN("mvf_dispconv_wgrad", mvf::dispconv_wgrad(xp, grad_y, grad_w, grad_b, workspace, B, C, H, W, (cudaStream_t)stream));
}

}  // extern "C"
