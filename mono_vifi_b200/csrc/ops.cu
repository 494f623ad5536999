// Stand-alone (unfused) kernels behind the reference's layers.py module API, so that the stock train.py
// runs on this library with zero edits.  Each is a single HBM-bound pass; the fused path (f1_*.cu) is the
// fast one.  Reference lines are cited per kernel.
#include "ops.cuh"

namespace mvf {

namespace {

constexpr int EW_THREADS = 256;
inline int ew_blocks(size_t n, int per_thread = 1) {
    size_t b = (n + (size_t)EW_THREADS * per_thread - 1) / ((size_t)EW_THREADS * per_thread);
    return (int)(b > 0 ? b : 1);
}

// ---- disp_to_depth (layers.py:16-25) -------------------------------------------------------------------
__global__ void disp_to_depth_fwd_k(const float* __restrict__ disp, float* __restrict__ sd_out,
                                    float* __restrict__ depth_out, size_t n, float min_disp, float range) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sd = __fadd_rn(min_disp, __fmul_rn(range, disp[i]));
    if (sd_out) sd_out[i] = sd;
    if (depth_out) depth_out[i] = __fdiv_rn(1.0f, sd);
}
__global__ void disp_to_depth_bwd_k(const float* __restrict__ disp, const float* __restrict__ g_sd,
                                    const float* __restrict__ g_depth, float* __restrict__ g_disp, size_t n,
                                    float min_disp, float range) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sd = min_disp + range * disp[i];
    float depth = 1.0f / sd;
    float g = 0.f;
    if (g_sd) g += g_sd[i] * range;
    if (g_depth) g -= g_depth[i] * range * depth * depth;
    g_disp[i] = g;
}

// ---- BackprojectDepth.forward (layers.py:192-197) ------------------------------------------------------
__global__ void backproject_fwd_k(const float* __restrict__ depth, const float* __restrict__ inv_K,
                                  float* __restrict__ out, int H, int W) {
    const int b = blockIdx.y;
    const size_t HW = (size_t)H * W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    int v = (int)(i / W), u = (int)(i - (size_t)v * W);
    float c[3];
    cam_ray(inv_K + 16 * b, (float)u, (float)v, c);
    float d = depth[b * HW + i];
    float* o = out + (size_t)b * 4 * HW + i;
    o[0] = __fmul_rn(d, c[0]);
    o[HW] = __fmul_rn(d, c[1]);
    o[2 * HW] = __fmul_rn(d, c[2]);
    o[3 * HW] = 1.0f;
}
__global__ void backproject_bwd_k(const float* __restrict__ g_out, const float* __restrict__ inv_K,
                                  float* __restrict__ g_depth, int H, int W) {
    const int b = blockIdx.y;
    const size_t HW = (size_t)H * W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    int v = (int)(i / W), u = (int)(i - (size_t)v * W);
    float c[3];
    cam_ray(inv_K + 16 * b, (float)u, (float)v, c);
    const float* g = g_out + (size_t)b * 4 * HW + i;
    g_depth[b * HW + i] = g[0] * c[0] + g[HW] * c[1] + g[2 * HW] * c[2];
}

// ---- Project3D.forward (layers.py:211-222), P = (K@T)[:, :3] computed by the caller in torch -------------
__global__ void project_fwd_k(const float* __restrict__ points, const float* __restrict__ P,
                              float* __restrict__ grid, int H, int W, float eps) {
    const int b = blockIdx.y;
    const size_t HW = (size_t)H * W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float* p = P + 12 * b;
    const float* X = points + (size_t)b * 4 * HW + i;
    float x0 = X[0], x1 = X[HW], x2 = X[2 * HW], x3 = X[3 * HW];
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(p[4 * r + 0], x0);
        acc = __fmaf_rn(p[4 * r + 1], x1, acc);
        acc = __fmaf_rn(p[4 * r + 2], x2, acc);
        acc = __fmaf_rn(p[4 * r + 3], x3, acc);
        c[r] = acc;
    }
    float z = __fadd_rn(c[2], eps);
    float x = __fdiv_rn(__fdiv_rn(c[0], z), (float)(W - 1));
    float y = __fdiv_rn(__fdiv_rn(c[1], z), (float)(H - 1));
    float2 o = make_float2(__fmul_rn(__fsub_rn(x, 0.5f), 2.0f), __fmul_rn(__fsub_rn(y, 0.5f), 2.0f));
    reinterpret_cast<float2*>(grid)[b * HW + i] = o;
}
// g_points [B,4,HW] and per-block partial g_P -> fixed-point atomics into accP [B,12]
__global__ void project_bwd_k(const float* __restrict__ points, const float* __restrict__ P,
                              const float* __restrict__ g_grid, float* __restrict__ g_points,
                              long long* __restrict__ accP, int H, int W, float eps) {
    const int b = blockIdx.y;
    const size_t HW = (size_t)H * W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float* p = P + 12 * b;
    float gP[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) gP[q] = 0.f;
    if (i < HW) {
        const float* X = points + (size_t)b * 4 * HW + i;
        float x[4] = {X[0], X[HW], X[2 * HW], X[3 * HW]};
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) c[r] = p[4 * r] * x[0] + p[4 * r + 1] * x[1] + p[4 * r + 2] * x[2] + p[4 * r + 3] * x[3];
        float rz = 1.0f / (c[2] + eps);
        float2 gg = reinterpret_cast<const float2*>(g_grid)[b * HW + i];
        float gx = gg.x * 2.0f / (float)(W - 1), gy = gg.y * 2.0f / (float)(H - 1);
        float gc[3] = {gx * rz, gy * rz, -(gx * c[0] + gy * c[1]) * rz * rz};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            g_points[(size_t)b * 4 * HW + j * HW + i] = gc[0] * p[j] + gc[1] * p[4 + j] + gc[2] * p[8 + j];
#pragma unroll
            for (int r = 0; r < 3; ++r) gP[4 * r + j] = gc[r] * x[j];
        }
    }
    __shared__ float red[EW_THREADS / 32][12];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        float v = warp_sum(gP[q]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double v = 0;
        for (int w = 0; w < EW_THREADS / 32; ++w) v += (double)red[w][threadIdx.x];
        atomicAdd(reinterpret_cast<unsigned long long*>(accP + 12 * b + threadIdx.x), (unsigned long long)to_fix(v));
    }
}
__global__ void fix_to_float_k(long long* __restrict__ acc, float* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (float)from_fix(acc[i]);
    acc[i] = 0;
}

// ---- SSIM.forward (layers.py:277-290) on [N,H,W] planes ------------------------------------------------
struct Win {
    float mx, my, vx, vy, vxy;
};
__device__ __forceinline__ Win window(const float* __restrict__ xp, const float* __restrict__ yp, int v, int u, int H,
                                      int W) {
    float sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int dv = -1; dv <= 1; ++dv) {
        const size_t ro = (size_t)reflect1(v + dv, H) * W;
#pragma unroll
        for (int du = -1; du <= 1; ++du) {
            size_t j = ro + reflect1(u + du, W);
            float a = __ldg(xp + j), t = __ldg(yp + j);
            sx += a;
            sy += t;
            sxx = fmaf(a, a, sxx);
            syy = fmaf(t, t, syy);
            sxy = fmaf(a, t, sxy);
        }
    }
    Win w;
    const float k9 = 1.0f / 9.0f;
    w.mx = sx * k9;
    w.my = sy * k9;
    w.vx = fmaf(sxx, k9, -w.mx * w.mx);
    w.vy = fmaf(syy, k9, -w.my * w.my);
    w.vxy = fmaf(sxy, k9, -w.mx * w.my);
    return w;
}
__global__ void ssim_fwd_k(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int H,
                           int W) {
    const size_t HW = (size_t)H * W;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= W || v >= H) return;
    const size_t po = (size_t)blockIdx.z * HW;
    Win w = window(x + po, y + po, v, u, H, W);
    const float C1 = 0.0001f, C2 = 0.0009f;
    float n = (2.0f * w.mx * w.my + C1) * (2.0f * w.vxy + C2);
    float d = (w.mx * w.mx + w.my * w.my + C1) * (w.vx + w.vy + C2);
    out[po + (size_t)v * W + u] = __saturatef((1.0f - n / d) * 0.5f);
}
// pass 1 of the backward: per-window adjoint coefficients (w.r.t. the FIRST argument) scaled by g_out
__global__ void ssim_bwd_coeff_k(const float* __restrict__ x, const float* __restrict__ y,
                                 const float* __restrict__ g_out, float* __restrict__ coef /*[3][N,H,W]*/, int H, int W,
                                 size_t plane_total) {
    const size_t HW = (size_t)H * W;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= W || v >= H) return;
    const size_t po = (size_t)blockIdx.z * HW, i = po + (size_t)v * W + u;
    Win w = window(x + po, y + po, v, u, H, W);
    const float C1 = 0.0001f, C2 = 0.0009f;
    float A = 2.0f * w.mx * w.my + C1, Bn = 2.0f * w.vxy + C2;
    float Cd = w.mx * w.mx + w.my * w.my + C1, D = w.vx + w.vy + C2;
    float inv = 1.0f / (Cd * D);
    float S = A * Bn * inv;
    float cf = (S >= -1.0f && S <= 1.0f) ? g_out[i] * (-0.5f / 9.0f) : 0.0f;
    coef[i] = cf * (2.0f * w.my * (Bn - A) - 2.0f * w.mx * S * (D - Cd)) * inv;
    coef[plane_total + i] = cf * (-2.0f * S * Cd * inv);
    coef[2 * plane_total + i] = cf * (2.0f * A * inv);
}
__device__ __forceinline__ float mlo(int q) { return q == 0 ? 0.0f : (q == 1 ? 2.0f : 1.0f); }
__device__ __forceinline__ float mhi(int q, int n) { return q == n - 1 ? 0.0f : (q == n - 2 ? 2.0f : 1.0f); }
// pass 2: 3x3 adjoint stencil with reflection multiplicity
__global__ void ssim_bwd_stencil_k(const float* __restrict__ x, const float* __restrict__ y,
                                   const float* __restrict__ coef, float* __restrict__ g_x, int H, int W,
                                   size_t plane_total) {
    const size_t HW = (size_t)H * W;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= W || v >= H) return;
    const size_t po = (size_t)blockIdx.z * HW, i = po + (size_t)v * W + u;
    float wy[3] = {mlo(v), 1.0f, mhi(v, H)}, wx[3] = {mlo(u), 1.0f, mhi(u, W)};
    float G0 = 0, G1 = 0, G2 = 0;
#pragma unroll
    for (int dv = 0; dv < 3; ++dv) {
        if (wy[dv] == 0.0f) continue;
#pragma unroll
        for (int du = 0; du < 3; ++du) {
            if (wx[du] == 0.0f) continue;
            size_t j = po + (size_t)(v + dv - 1) * W + (u + du - 1);
            float wgt = wy[dv] * wx[du];
            G0 = fmaf(wgt, __ldg(coef + j), G0);
            G1 = fmaf(wgt, __ldg(coef + plane_total + j), G1);
            G2 = fmaf(wgt, __ldg(coef + 2 * plane_total + j), G2);
        }
    }
    g_x[i] = G0 + x[i] * G1 + y[i] * G2;
}

// ---- get_smooth_loss (layers.py:231-242) ---------------------------------------------------------------
// acc[0] += sum_x, acc[1] += sum_y (fixed point); finalised by smooth_final_k
__global__ void smooth_fwd_k(const float* __restrict__ disp, const float* __restrict__ img, long long* __restrict__ acc,
                             int H, int W) {
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    float sx = 0.f, sy = 0.f;
    if (u < W && v < H) {
        const float* d = disp + b * HW;
        const float* im = img + (size_t)b * 3 * HW;
        size_t i = (size_t)v * W + u;
        float d0 = d[i], i0 = im[i], i1 = im[HW + i], i2 = im[2 * HW + i];
        if (u + 1 < W) {
            float gi = fabsf(i0 - im[i + 1]) + fabsf(i1 - im[HW + i + 1]) + fabsf(i2 - im[2 * HW + i + 1]);
            sx = fabsf(d0 - d[i + 1]) * __expf(-(gi / 3.0f));
        }
        if (v + 1 < H) {
            float gi = fabsf(i0 - im[i + W]) + fabsf(i1 - im[HW + i + W]) + fabsf(i2 - im[2 * HW + i + W]);
            sy = fabsf(d0 - d[i + W]) * __expf(-(gi / 3.0f));
        }
    }
    __shared__ float red[8][2];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    if ((tid & 31) == 0) { red[tid >> 5][0] = sx; red[tid >> 5][1] = sy; }
    __syncthreads();
    if (tid < 2) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += (double)red[w][tid];
        atomicAdd(reinterpret_cast<unsigned long long*>(acc + tid), (unsigned long long)to_fix(s));
    }
}
__global__ void smooth_final_k(long long* __restrict__ acc, float* __restrict__ out, int B, int H, int W) {
    double sx = from_fix(acc[0]), sy = from_fix(acc[1]);
    acc[0] = 0;
    acc[1] = 0;
    out[0] = (float)(sx / ((double)B * H * (W - 1)) + sy / ((double)B * (H - 1) * W));
}
__global__ void smooth_bwd_k(const float* __restrict__ disp, const float* __restrict__ img,
                             const float* __restrict__ gout, float* __restrict__ g_disp, int B, int H, int W) {
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= W || v >= H) return;
    const float go = gout ? gout[0] : 1.0f;
    const float cx = go / (float)((double)B * H * (W - 1)), cy = go / (float)((double)B * (H - 1) * W);
    const float* d = disp + b * HW;
    const float* im = img + (size_t)b * 3 * HW;
    size_t i = (size_t)v * W + u;
    float d0 = d[i], i0 = im[i], i1 = im[HW + i], i2 = im[2 * HW + i];
    float g = 0.f;
    auto edge = [&](size_t j) {
        float gi = fabsf(i0 - im[j]) + fabsf(i1 - im[HW + j]) + fabsf(i2 - im[2 * HW + j]);
        return __expf(-(gi / 3.0f));
    };
    auto sgn = [](float a) { return a > 0.f ? 1.0f : (a < 0.f ? -1.0f : 0.0f); };
    if (u + 1 < W) g += cx * sgn(d0 - d[i + 1]) * edge(i + 1);
    if (u >= 1) g -= cx * sgn(d[i - 1] - d0) * edge(i - 1);
    if (v + 1 < H) g += cy * sgn(d0 - d[i + W]) * edge(i + W);
    if (v >= 1) g -= cy * sgn(d[i - W] - d0) * edge(i - W);
    g_disp[b * HW + i] = g;
}

// ---- compute_SI_log_depth_loss (train.py:924-941) -------------------------------------------------------
// acc [B][3] = {n, sum d, sum d^2} in fixed point
__global__ void si_log_fwd_k(const float* __restrict__ pred, const float* __restrict__ target,
                             const float* __restrict__ mask, long long* __restrict__ acc, size_t HW) {
    const int b = blockIdx.y;
    float n = 0.f, s1 = 0.f, s2 = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x) {
        float m = mask ? mask[b * HW + i] : 1.0f;
        float ld = __logf(pred[b * HW + i] + 1e-7f) * m - __logf(target[b * HW + i] + 1e-7f) * m;
        n += m;
        s1 += ld;
        s2 = fmaf(ld, ld, s2);
    }
    __shared__ float red[EW_THREADS / 32][3];
    n = warp_sum(n);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = n; red[threadIdx.x >> 5][1] = s1; red[threadIdx.x >> 5][2] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0;
        for (int w = 0; w < EW_THREADS / 32; ++w) s += (double)red[w][threadIdx.x];
        atomicAdd(reinterpret_cast<unsigned long long*>(acc + 3 * b + threadIdx.x), (unsigned long long)to_fix(s));
    }
}
// loss = mean_b( s2/n - beta*s1^2/n^2 ); stats[b] = {n, s1} kept for the backward
__global__ void si_log_final_k(long long* __restrict__ acc, float* __restrict__ loss, float* __restrict__ stats, int B,
                               float beta) {
    double total = 0;
    for (int b = 0; b < B; ++b) {
        double n = from_fix(acc[3 * b]) + 1e-8, s1 = from_fix(acc[3 * b + 1]), s2 = from_fix(acc[3 * b + 2]);
        acc[3 * b] = acc[3 * b + 1] = acc[3 * b + 2] = 0;
        total += s2 / n - (double)beta * s1 * s1 / (n * n);
        stats[2 * b] = (float)n;
        stats[2 * b + 1] = (float)s1;
    }
    loss[0] = (float)(total / B);
}
__global__ void si_log_bwd_k(const float* __restrict__ pred, const float* __restrict__ target,
                             const float* __restrict__ mask, const float* __restrict__ stats,
                             const float* __restrict__ gout, float* __restrict__ g_pred, float* __restrict__ g_target,
                             size_t HW, int B, float beta) {
    const int b = blockIdx.y;
    const float n = stats[2 * b], s1 = stats[2 * b + 1];
    const float go = (gout ? gout[0] : 1.0f) / (float)B;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x) {
        float m = mask ? mask[b * HW + i] : 1.0f;
        float p = pred[b * HW + i] + 1e-7f, t = target[b * HW + i] + 1e-7f;
        float ld = __logf(p) * m - __logf(t) * m;
        float g = go * (2.0f * ld / n - 2.0f * beta * s1 / (n * n)) * m;
        if (g_pred) g_pred[b * HW + i] = g / p;
        if (g_target) g_target[b * HW + i] = -g / t;
    }
}

}  // namespace

#define LAUNCH_CHECK() return cudaGetLastError()

cudaError_t disp_to_depth_fwd(const float* disp, float* sd, float* depth, size_t n, float min_disp, float range,
                              cudaStream_t st) {
    disp_to_depth_fwd_k<<<ew_blocks(n), EW_THREADS, 0, st>>>(disp, sd, depth, n, min_disp, range);
    LAUNCH_CHECK();
}
cudaError_t disp_to_depth_bwd(const float* disp, const float* g_sd, const float* g_depth, float* g_disp, size_t n,
                              float min_disp, float range, cudaStream_t st) {
    disp_to_depth_bwd_k<<<ew_blocks(n), EW_THREADS, 0, st>>>(disp, g_sd, g_depth, g_disp, n, min_disp, range);
    LAUNCH_CHECK();
}
cudaError_t backproject_fwd(const float* depth, const float* inv_K, float* out, int B, int H, int W, cudaStream_t st) {
    dim3 grid(ew_blocks((size_t)H * W), B);
    backproject_fwd_k<<<grid, EW_THREADS, 0, st>>>(depth, inv_K, out, H, W);
    LAUNCH_CHECK();
}
cudaError_t backproject_bwd(const float* g_out, const float* inv_K, float* g_depth, int B, int H, int W,
                            cudaStream_t st) {
    dim3 grid(ew_blocks((size_t)H * W), B);
    backproject_bwd_k<<<grid, EW_THREADS, 0, st>>>(g_out, inv_K, g_depth, H, W);
    LAUNCH_CHECK();
}
cudaError_t project_fwd(const float* points, const float* P, float* grid_out, int B, int H, int W, float eps,
                        cudaStream_t st) {
    dim3 grid(ew_blocks((size_t)H * W), B);
    project_fwd_k<<<grid, EW_THREADS, 0, st>>>(points, P, grid_out, H, W, eps);
    LAUNCH_CHECK();
}
cudaError_t project_bwd(const float* points, const float* P, const float* g_grid, float* g_points, float* g_P,
                        long long* acc, int B, int H, int W, float eps, cudaStream_t st) {
    dim3 grid(ew_blocks((size_t)H * W), B);
    project_bwd_k<<<grid, EW_THREADS, 0, st>>>(points, P, g_grid, g_points, acc, H, W, eps);
    fix_to_float_k<<<(12 * B + 127) / 128, 128, 0, st>>>(acc, g_P, 12 * B);
    LAUNCH_CHECK();
}
cudaError_t ssim_fwd(const float* x, const float* y, float* out, int N, int H, int W, cudaStream_t st) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, N);
    ssim_fwd_k<<<grid, blk, 0, st>>>(x, y, out, H, W);
    LAUNCH_CHECK();
}
cudaError_t ssim_bwd(const float* x, const float* y, const float* g_out, float* g_x, float* coef_ws, int N, int H,
                     int W, cudaStream_t st) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, N);
    size_t total = (size_t)N * H * W;
    ssim_bwd_coeff_k<<<grid, blk, 0, st>>>(x, y, g_out, coef_ws, H, W, total);
    ssim_bwd_stencil_k<<<grid, blk, 0, st>>>(x, y, coef_ws, g_x, H, W, total);
    LAUNCH_CHECK();
}
cudaError_t smooth_fwd(const float* disp, const float* img, float* out, long long* acc, int B, int H, int W,
                       cudaStream_t st) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, B);
    smooth_fwd_k<<<grid, blk, 0, st>>>(disp, img, acc, H, W);
    smooth_final_k<<<1, 1, 0, st>>>(acc, out, B, H, W);
    LAUNCH_CHECK();
}
cudaError_t smooth_bwd(const float* disp, const float* img, const float* gout, float* g_disp, int B, int H, int W,
                       cudaStream_t st) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, B);
    smooth_bwd_k<<<grid, blk, 0, st>>>(disp, img, gout, g_disp, B, H, W);
    LAUNCH_CHECK();
}
cudaError_t si_log_fwd(const float* pred, const float* target, const float* mask, float* loss, float* stats,
                       long long* acc, int B, size_t HW, float beta, cudaStream_t st) {
    int bx = (int)((HW + EW_THREADS * 4 - 1) / (EW_THREADS * 4));
    if (bx > 148) bx = 148;
    dim3 grid(bx, B);
    si_log_fwd_k<<<grid, EW_THREADS, 0, st>>>(pred, target, mask, acc, HW);
    si_log_final_k<<<1, 1, 0, st>>>(acc, loss, stats, B, beta);
    LAUNCH_CHECK();
}
cudaError_t si_log_bwd(const float* pred, const float* target, const float* mask, const float* stats,
                       const float* gout, float* g_pred, float* g_target, int B, size_t HW, float beta,
                       cudaStream_t st) {
    int bx = (int)((HW + EW_THREADS * 4 - 1) / (EW_THREADS * 4));
    dim3 grid(bx, B);
    si_log_bwd_k<<<grid, EW_THREADS, 0, st>>>(pred, target, mask, stats, gout, g_pred, g_target, HW, B, beta);
    LAUNCH_CHECK();
}

}  // namespace mvf
