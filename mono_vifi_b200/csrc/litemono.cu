// HBM-bound kernels of the Lite-Mono blocks (networks/LiteMono.py), channels-last, float4 over channels:
//   dwconv3x3   depth-wise dilated 3x3 convolution, stride 1, zero padding = dilation (CDilated, LiteMono.py:140-155, used by
//               DilatedConv :157-201).  Forward; data gradient = the same kernel with the taps mirrored; weight gradient = a
//               per-channel reduction over all pixels (per-CTA partials added in a fixed order: bitwise reproducible).
//               cuDNN's grouped-convolution kernels are what the reference runs here (9 MAC per element: pure data movement).
//   gelu        exact (erf) GELU of nn.GELU(), forward and backward from the saved input.
//   layernorm   LayerNorm over the channel dimension of channels-last tokens (LiteMono.py:93-121, eps inside the sqrt), one warp
//               per pixel; backward: grad_x per pixel, grad_weight / grad_bias through fixed-order per-CTA partials.
// Dense channels-last tensors [P pixels][C], C % 4 == 0.
#include "litemono.cuh"

#include "pdl.cuh"

namespace mvf {
namespace {

constexpr int NT = 256;

inline int grid_for(long long total, int cap_waves = 16) {
    long long g = (total + NT - 1) / NT;
    const long long cap = 148LL * cap_waves;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

// ---- depth-wise dilated 3x3 ----------------------------------------------------------------------------------------------
// w_t: taps-major copy of the filter, [9][C] (tap = kh * 3 + kw); flip = 1 evaluates the data gradient (tap 8 - t)
__global__ void dwconv3x3_kernel(const float4* __restrict__ x, const float4* __restrict__ w_t, const float4* __restrict__ bias,
                                 float4* __restrict__ y, int B, int C4, int H, int W, int dil, int flip) {
    pdl_sync();
    const long long total = (long long)B * H * W * C4;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int px = (int)(r % W);
        r /= W;
        const int py = (int)(r % H), b = (int)(r / H);
        float4 acc = bias ? __ldg(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* img = x + (long long)b * H * W * C4 + c;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int sy = py + (kh - 1) * dil;
            if (sy < 0 || sy >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int sx = px + (kw - 1) * dil;
                if (sx < 0 || sx >= W) continue;
                const int t = kh * 3 + kw;
                acc = fma4(__ldg(w_t + (flip ? 8 - t : t) * C4 + c), __ldg(img + ((long long)sy * W + sx) * C4), acc);
            }
        }
        y[i] = acc;
    }
}

// weight gradient: partial[block][9][C]; threads = (row, channel group), every CTA walks a contiguous range of pixels
__global__ void __launch_bounds__(NT) dwconv3x3_wgrad_partial_kernel(const float4* __restrict__ x, const float4* __restrict__ gy,
                                                                     float* __restrict__ partial, int B, int C4, int H, int W, int dil,
                                                                     long long px_per_block) {
    pdl_sync();
    extern __shared__ float4 red[];   // [rows][9][C4]
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    const int c = threadIdx.x % C4, r = threadIdx.x / C4;
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long P = (long long)B * H * W;
    const long long p0 = blockIdx.x * px_per_block, p1 = min(P, p0 + px_per_block);
    if (r < rows) {
        for (long long p = p0 + r; p < p1; p += rows) {
            const int px = (int)(p % W);
            const long long q = p / W;
            const int py = (int)(q % H);
            const long long img = (q / H) * H * W;
            const float4 g = __ldg(gy + p * C4 + c);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int sy = py + (kh - 1) * dil;
                if (sy < 0 || sy >= H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int sx = px + (kw - 1) * dil;
                    if (sx < 0 || sx >= W) continue;
                    acc[kh * 3 + kw] = fma4(g, __ldg(x + (img + (long long)sy * W + sx) * C4 + c), acc[kh * 3 + kw]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) red[(r * 9 + t) * C4 + c] = acc[t];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 9 * C4; j += NT) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < rows; ++q) {
            const float4 u = red[q * 9 * C4 + j];
            s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w;
        }
        reinterpret_cast<float4*>(partial)[(size_t)blockIdx.x * 9 * C4 + j] = s;
    }
}
// gw[c][t] = sum over blocks (fixed order) of partial[block][t][c]; one thread per (t, c)
__global__ void dwconv3x3_wgrad_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ gw) {
    pdl_sync();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 9 * C) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 9 * C + j];
    const int t = j / C, c = j % C;
    gw[c * 9 + t] = s;
}

// ---- GELU ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu1(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu1(float v) {
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
    return cdf + v * 0.3989422804014327f * __expf(-0.5f * v * v);
}
__global__ void gelu_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        const float4 v = __ldg(x + i);
        y[i] = make_float4(gelu1(v.x), gelu1(v.y), gelu1(v.z), gelu1(v.w));
    }
}
__global__ void gelu_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, float4* __restrict__ gx, long long n4) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        const float4 v = __ldg(x + i), g = __ldg(gy + i);
        gx[i] = make_float4(g.x * dgelu1(v.x), g.y * dgelu1(v.y), g.z * dgelu1(v.z), g.w * dgelu1(v.w));
    }
}

// ---- LayerNorm over channels ---------------------------------------------------------------------------------------------
// one warp per pixel; lane l holds channel groups l, l + 32, ... (C4 <= 32 * LN_MAXG)
constexpr int LN_MAXG = 4;   // C <= 512
__global__ void layernorm_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ w, const float4* __restrict__ b,
                                     float4* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, long long P, int C4,
                                     float eps) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)NT + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * NT) >> 5;
    const float invC = 1.f / (float)(4 * C4);
    for (long long p = warp0; p < P; p += nwarps) {
        float4 v[LN_MAXG];
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < LN_MAXG; ++g) {
            const int c = lane + 32 * g;
            v[g] = c < C4 ? __ldg(x + p * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[g].x + v[g].y) + (v[g].z + v[g].w);
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mu = s * invC;
        float q = 0.f;
#pragma unroll
        for (int g = 0; g < LN_MAXG; ++g) {
            if (lane + 32 * g < C4) {
                const float a = v[g].x - mu, bb = v[g].y - mu, cc = v[g].z - mu, d = v[g].w - mu;
                q += (a * a + bb * bb) + (cc * cc + d * d);
            }
        }
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rs = rsqrtf(q * invC + eps);
        if (lane == 0) {
            mean[p] = mu;
            rstd[p] = rs;
        }
#pragma unroll
        for (int g = 0; g < LN_MAXG; ++g) {
            const int c = lane + 32 * g;
            if (c < C4) {
                const float4 ww = __ldg(w + c), bv = __ldg(b + c);
                y[p * C4 + c] = make_float4(fmaf((v[g].x - mu) * rs, ww.x, bv.x), fmaf((v[g].y - mu) * rs, ww.y, bv.y),
                                            fmaf((v[g].z - mu) * rs, ww.z, bv.z), fmaf((v[g].w - mu) * rs, ww.w, bv.w));
            }
        }
    }
}

// grad_x per pixel (one warp) + per-CTA partial sums of grad_w / grad_b: partial[block][2][C]
__global__ void __launch_bounds__(NT) layernorm_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gy,
                                                           const float4* __restrict__ w, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float4* __restrict__ gx,
                                                           float* __restrict__ partial, long long P, int C4, long long px_per_block) {
    pdl_sync();
    extern __shared__ float4 red[];   // [8 warps][2][C4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float invC = 1.f / (float)(4 * C4);
    float4 gw[LN_MAXG], gb[LN_MAXG];
#pragma unroll
    for (int g = 0; g < LN_MAXG; ++g) gw[g] = gb[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long p0 = blockIdx.x * px_per_block, p1 = min(P, p0 + px_per_block);
    for (long long p = p0 + warp; p < p1; p += NT / 32) {
        const float mu = __ldg(mean + p), rs = __ldg(rstd + p);
        float4 xh[LN_MAXG], gg[LN_MAXG];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int g = 0; g < LN_MAXG; ++g) {
            const int c = lane + 32 * g;
            if (c < C4) {
                const float4 v = __ldg(x + p * C4 + c), dy = __ldg(gy + p * C4 + c), ww = __ldg(w + c);
                xh[g] = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
                gg[g] = make_float4(dy.x * ww.x, dy.y * ww.y, dy.z * ww.z, dy.w * ww.w);
                s1 += (gg[g].x + gg[g].y) + (gg[g].z + gg[g].w);
                s2 += (gg[g].x * xh[g].x + gg[g].y * xh[g].y) + (gg[g].z * xh[g].z + gg[g].w * xh[g].w);
                gw[g].x = fmaf(dy.x, xh[g].x, gw[g].x); gw[g].y = fmaf(dy.y, xh[g].y, gw[g].y);
                gw[g].z = fmaf(dy.z, xh[g].z, gw[g].z); gw[g].w = fmaf(dy.w, xh[g].w, gw[g].w);
                gb[g].x += dy.x; gb[g].y += dy.y; gb[g].z += dy.z; gb[g].w += dy.w;
            }
        }
        for (int o = 16; o; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 * invC, m2 = s2 * invC;
#pragma unroll
        for (int g = 0; g < LN_MAXG; ++g) {
            const int c = lane + 32 * g;
            if (c < C4)
                gx[p * C4 + c] = make_float4(rs * (gg[g].x - m1 - xh[g].x * m2), rs * (gg[g].y - m1 - xh[g].y * m2),
                                             rs * (gg[g].z - m1 - xh[g].z * m2), rs * (gg[g].w - m1 - xh[g].w * m2));
        }
    }
#pragma unroll
    for (int g = 0; g < LN_MAXG; ++g) {
        const int c = lane + 32 * g;
        if (c < C4) {
            red[(warp * 2 + 0) * C4 + c] = gw[g];
            red[(warp * 2 + 1) * C4 + c] = gb[g];
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * C4; j += NT) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < NT / 32; ++q) {
            const float4 u = red[q * 2 * C4 + j];
            s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w;
        }
        reinterpret_cast<float4*>(partial)[(size_t)blockIdx.x * 2 * C4 + j] = s;
    }
}
__global__ void layernorm_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ gw,
                                              float* __restrict__ gb) {
    pdl_sync();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * C) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 2 * C + j];
    (j < C ? gw : gb)[j % C] = s;
}

int reduce_blocks(long long P) {
    long long nb = (P + 255) / 256;   // at least ~256 pixels per CTA
    return (int)(nb < 1 ? 1 : (nb > 148 * 4 ? 148 * 4 : nb));
}

}  // namespace

cudaError_t dwconv3x3_fwd(const float* x, const float* w_taps, const float* bias, float* y, int B, int C, int H, int W, int dil, int flip,
                          cudaStream_t st) {
    const long long total = (long long)B * H * W * (C / 4);
    return launch_pdl(dwconv3x3_kernel, dim3(grid_for(total)), dim3(NT), 0, st, (const float4*)x, (const float4*)w_taps, (const float4*)bias,
                      (float4*)y, B, C / 4, H, W, dil, flip);
}

size_t dwconv3x3_wgrad_workspace_floats(long long P, int C) { return (size_t)reduce_blocks(P) * 9 * C; }

cudaError_t dwconv3x3_wgrad(const float* x, const float* gy, float* gw, float* workspace, int B, int C, int H, int W, int dil,
                            cudaStream_t st) {
    const long long P = (long long)B * H * W;
    const int nb = reduce_blocks(P), C4 = C / 4;
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    const long long per = (P + nb - 1) / nb;
    const size_t smem = (size_t)rows * 9 * C4 * sizeof(float4);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dwconv3x3_wgrad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaError_t e = launch_pdl(dwconv3x3_wgrad_partial_kernel, dim3(nb), dim3(NT), smem, st, (const float4*)x, (const float4*)gy, workspace,
                               B, C4, H, W, dil, per);
    if (e != cudaSuccess) return e;
    return launch_pdl(dwconv3x3_wgrad_finalize_kernel, dim3((9 * C + 127) / 128), dim3(128), 0, st, (const float*)workspace, nb, C, gw);
}

cudaError_t gelu_fwd(const float* x, float* y, long long n, cudaStream_t st) {
    return launch_pdl(gelu_fwd_kernel, dim3(grid_for(n / 4)), dim3(NT), 0, st, (const float4*)x, (float4*)y, n / 4);
}
cudaError_t gelu_bwd(const float* x, const float* gy, float* gx, long long n, cudaStream_t st) {
    return launch_pdl(gelu_bwd_kernel, dim3(grid_for(n / 4)), dim3(NT), 0, st, (const float4*)x, (const float4*)gy, (float4*)gx, n / 4);
}

cudaError_t layernorm_cl_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, long long P, int C,
                             float eps, cudaStream_t st) {
    return launch_pdl(layernorm_fwd_kernel, dim3(grid_for(P * 32)), dim3(NT), 0, st, (const float4*)x, (const float4*)w, (const float4*)b,
                      (float4*)y, mean, rstd, P, C / 4, eps);
}
size_t layernorm_bwd_workspace_floats(long long P, int C) { return (size_t)reduce_blocks(P) * 2 * C; }
cudaError_t layernorm_cl_bwd(const float* x, const float* gy, const float* w, const float* mean, const float* rstd, float* gx, float* gw,
                             float* gb, float* workspace, long long P, int C, cudaStream_t st) {
    const int nb = reduce_blocks(P), C4 = C / 4;
    const long long per = (P + nb - 1) / nb;
    cudaError_t e = launch_pdl(layernorm_bwd_kernel, dim3(nb), dim3(NT), (size_t)(NT / 32) * 2 * C4 * sizeof(float4), st, (const float4*)x,
                               (const float4*)gy, (const float4*)w, mean, rstd, (float4*)gx, workspace, P, C4, per);
    if (e != cudaSuccess) return e;
    return launch_pdl(layernorm_bwd_finalize_kernel, dim3((2 * C + 127) / 128), dim3(128), 0, st, (const float*)workspace, nb, C, gw, gb);
}

}  // namespace mvf
