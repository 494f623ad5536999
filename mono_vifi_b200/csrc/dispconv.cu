// The disparity head: 3x3 convolution with ONE output channel on a (reflection-)padded channels-last input (Conv3x3(num_ch_dec[s], 1),
// networks/monodepth2.py:76-77, LiteMono.py:468-469, DHRNet.py; applied at full resolution).  On the tensor cores a 1-channel output
// is an N = 16 tile with 15 zero columns: 119 us forward, 113 us data gradient, 169 us weight gradient at B12 192x640 for 0.4 GFLOP
// each -- all on the critical path of the step.  The layer is pure data movement (144 MAC per output pixel, 64 B read per pixel), so
// these are plain HBM-bound kernels:
//   fwd   y[b,y,x]        = bias + sum_{kh,kw,c} w[c,kh,kw] * xp[b, y+kh, x+kw, c]
//   dgrad gxp[b,py,px,c]  = sum_{kh,kw} w[c,kh,kw] * gy[b, py-kh, px-kw]            (gather form, zero outside)
//   wgrad gw[c,kh,kw]     = sum_{b,y,x} gy[b,y,x] * xp[b, y+kh, x+kw, c],  gb = sum gy   (per-CTA partials, fixed-order finalize)
// xp: dense channels-last [B, H+2, W+2, C], C % 4 == 0, C <= 64; y / gy: [B, H, W].
#include "dispconv.cuh"

#include "pdl.cuh"

namespace mvf {
namespace {

constexpr int NT = 256;
constexpr int RPB = 8;   // image rows per CTA (fwd, dgrad): the staged weights and the index arithmetic are paid once per 8 rows

// Work decomposition (all three kernels): blockIdx.y (or a block-strided loop) walks image rows, so that the only divisions are one
// per row; the 4-channel groups of a pixel sit on neighbouring lanes (tpp = C/4 rounded up to a power of two), so a warp's load of one
// tap is one contiguous run of 32 / tpp pixels (the first version -- one thread per pixel walking its 64 B with the neighbouring
// lane 64 B further on -- cost 4x the L1 wavefronts and 64-bit div/mod chains per element: 126 / 80 / 139 us at B12 192x640).
__device__ __forceinline__ void stage_weights(float4* ws, const float* __restrict__ w, int C4) {
    for (int i = threadIdx.x; i < 9 * C4; i += NT) {   // [9][C4] tap-major
        const int t = i / C4, c = i - t * C4;
        ws[i] = make_float4(w[(4 * c + 0) * 9 + t], w[(4 * c + 1) * 9 + t], w[(4 * c + 2) * 9 + t], w[(4 * c + 3) * 9 + t]);
    }
}

// fwd: a thread owns (pixel column x, channel group c) and walks RPB image rows downwards with a 3-row register window: per output row it
// loads the 3 float4 of the new bottom row only (ncu of the 9-loads-per-row version: 38 M warp-instructions = 26 per pixel, issue-bound
// at 76 us with 66 registers; the window needs a third of the loads and of the address arithmetic).
__global__ void __launch_bounds__(NT) dispconv_fwd_kernel(const float4* __restrict__ xp, const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ y, int rows, int C4, int H, int W, int tpp_log2) {
    pdl_sync();
    extern __shared__ float4 ws[];
    stage_weights(ws, w, C4);
    __syncthreads();
    const int t = blockIdx.x * NT + threadIdx.x;
    const int x = t >> tpp_log2, c = t & ((1 << tpp_log2) - 1);
    const int Wp = W + 2;
    const float b0 = bias ? bias[0] : 0.f;
    const bool live = x < W && c < C4;
    const float4* wq = ws + (live ? c : 0);
    const int row0 = blockIdx.y * RPB, row_end = min(rows, row0 + RPB);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v0[3] = {z, z, z}, v1[3] = {z, z, z}, v2[3];
    const float4* src = nullptr;          // padded row yy + 2 of the current image at column x
    for (int row = row0; row < row_end; ++row) {   // row = b * H + yy
        const int b = row / H, yy = row - b * H;
        if (row == row0 || yy == 0) {     // (re)fill the window at the top of the block's range / of an image
            if (live) {
                const float4* top = xp + (((long long)b * (H + 2) + yy) * Wp + x) * C4 + c;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    v0[kw] = __ldg(top + kw * C4);
                    v1[kw] = __ldg(top + ((long long)Wp + kw) * C4);
                }
                src = top + 2LL * Wp * C4;
            }
        }
        float acc = 0.f;
        if (live) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) v2[kw] = __ldg(src + kw * C4);
            src += (long long)Wp * C4;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float4 q0 = wq[(0 + kw) * C4], q1 = wq[(3 + kw) * C4], q2 = wq[(6 + kw) * C4];
                a0 = fmaf(v0[kw].x, q0.x, a0); a1 = fmaf(v0[kw].y, q0.y, a1); a2 = fmaf(v0[kw].z, q0.z, a2); a3 = fmaf(v0[kw].w, q0.w, a3);
                a0 = fmaf(v1[kw].x, q1.x, a0); a1 = fmaf(v1[kw].y, q1.y, a1); a2 = fmaf(v1[kw].z, q1.z, a2); a3 = fmaf(v1[kw].w, q1.w, a3);
                a0 = fmaf(v2[kw].x, q2.x, a0); a1 = fmaf(v2[kw].y, q2.y, a1); a2 = fmaf(v2[kw].z, q2.z, a2); a3 = fmaf(v2[kw].w, q2.w, a3);
            }
            acc = (a0 + a1) + (a2 + a3);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                v0[kw] = v1[kw];
                v1[kw] = v2[kw];
            }
        }
        for (int o = 1; o < (1 << tpp_log2); o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // fixed butterfly over the pixel's lanes
        if (x < W && c == 0) y[(long long)row * W + x] = acc + b0;
    }
}

// dgrad: thread = (padded pixel column px, channel group c), walking RPB padded rows with a 3 x 3 window of grad_y scalars (zero outside)
__global__ void __launch_bounds__(NT) dispconv_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ w, float4* __restrict__ gxp,
                                                            int rows, int C4, int H, int W) {
    pdl_sync();
    extern __shared__ float4 ws[];
    stage_weights(ws, w, C4);
    __syncthreads();
    const int Hp = H + 2, Wp = W + 2;
    const int t = blockIdx.x * NT + threadIdx.x;
    if (t >= Wp * C4) return;
    const int px = t / C4, c = t - px * C4;
    const float4* wq = ws + c;
    const bool okc[3] = {px < W, px >= 1 && px - 1 < W, px >= 2};   // column px - kw inside [0, W)
    const int row0 = blockIdx.y * RPB, row_end = min(rows, row0 + RPB);
    float g0[3] = {0.f, 0.f, 0.f}, g1[3] = {0.f, 0.f, 0.f}, g2[3];   // grad_y rows py-2, py-1, py at columns px - kw
    for (int row = row0; row < row_end; ++row) {   // row = b * Hp + py
        const int b = row / Hp, py = row - b * Hp;
        const float* g = gy + (long long)b * H * W + px;
        if (row == row0 || py == 0) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                g0[kw] = (py >= 2 && py - 2 < H && okc[kw]) ? __ldg(g + (long long)(py - 2) * W - kw) : 0.f;
                g1[kw] = (py >= 1 && py - 1 < H && okc[kw]) ? __ldg(g + (long long)(py - 1) * W - kw) : 0.f;
            }
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) g2[kw] = (py < H && okc[kw]) ? __ldg(g + (long long)py * W - kw) : 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            // output row oy = py - kh: kh = 0 reads g2, kh = 1 reads g1, kh = 2 reads g0
            const float4 q0 = wq[(0 + kw) * C4], q1 = wq[(3 + kw) * C4], q2 = wq[(6 + kw) * C4];
            acc.x = fmaf(g2[kw], q0.x, acc.x); acc.y = fmaf(g2[kw], q0.y, acc.y); acc.z = fmaf(g2[kw], q0.z, acc.z); acc.w = fmaf(g2[kw], q0.w, acc.w);
            acc.x = fmaf(g1[kw], q1.x, acc.x); acc.y = fmaf(g1[kw], q1.y, acc.y); acc.z = fmaf(g1[kw], q1.z, acc.z); acc.w = fmaf(g1[kw], q1.w, acc.w);
            acc.x = fmaf(g0[kw], q2.x, acc.x); acc.y = fmaf(g0[kw], q2.y, acc.y); acc.z = fmaf(g0[kw], q2.z, acc.z); acc.w = fmaf(g0[kw], q2.w, acc.w);
        }
        gxp[(long long)row * Wp * C4 + t] = acc;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            g0[kw] = g1[kw];
            g1[kw] = g2[kw];
        }
    }
}

// partial[block][9][C] (+ the bias gradient at [9*C]); threads = (pixel lane l, channel group c); a CTA walks rows blockIdx.x, += gridDim.x
__global__ void __launch_bounds__(NT) dispconv_wgrad_partial_kernel(const float4* __restrict__ xp, const float* __restrict__ gy,
                                                                    float* __restrict__ partial, int rows, int C4, int H, int W) {
    pdl_sync();
    extern __shared__ float4 red[];   // [lanes][9][C4] + bias column
    const int lanes = NT / C4;
    const int l = threadIdx.x / C4, c = threadIdx.x - l * C4;
    const int Wp = W + 2;
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    float gsum = 0.f;
    if (l < lanes) {
        for (int row = blockIdx.x; row < rows; row += gridDim.x) {
            const int b = row / H, yy = row - b * H;
            const float* grow = gy + (long long)row * W;
            const float4* xrow = xp + (((long long)b * (H + 2) + yy) * Wp) * C4 + c;
            for (int x = l; x < W; x += lanes) {
                const float g = __ldg(grow + x);
                if (c == 0) gsum += g;
                const float4* base = xrow + (long long)x * C4;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 v = __ldg(base + ((long long)kh * Wp + kw) * C4);
                        float4& a = acc[kh * 3 + kw];
                        a.x = fmaf(g, v.x, a.x); a.y = fmaf(g, v.y, a.y); a.z = fmaf(g, v.z, a.z); a.w = fmaf(g, v.w, a.w);
                    }
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) red[(l * 9 + t) * C4 + c] = acc[t];
    }
    float* gred = reinterpret_cast<float*>(red + (size_t)lanes * 9 * C4);
    if (l < lanes && c == 0) gred[l] = gsum;
    __syncthreads();
    const int C = 4 * C4;
    for (int j = threadIdx.x; j < 9 * C4; j += NT) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < lanes; ++q) {
            const float4 u = red[q * 9 * C4 + j];
            s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w;
        }
        reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * (9 * C + 4))[j] = s;
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int q = 0; q < lanes; ++q) s += gred[q];
        partial[(size_t)blockIdx.x * (9 * C + 4) + 9 * C] = s;
    }
}

// sums the per-CTA partials in a fixed order: thread (jl, bl) of a 32 x 32 CTA adds blocks bl, bl + 32, ... of output j0 + jl (for one
// block the 32 outputs are a coalesced 128-byte read), the 32 lanes of an output meet in shared memory
__global__ void __launch_bounds__(1024) dispconv_wgrad_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ gw,
                                                                       float* __restrict__ gb) {
    pdl_sync();
    __shared__ double red[32][33];
    const int jl = threadIdx.x & 31, bl = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + jl;
    double s = 0.0;
    if (j <= 9 * C)
        for (int b = bl; b < nblocks; b += 32) s += (double)partial[(size_t)b * (9 * C + 4) + j];
    red[bl][jl] = s;
    __syncthreads();
    if (bl != 0 || j > 9 * C) return;
    for (int q = 1; q < 32; ++q) s += red[q][jl];
    if (j == 9 * C) {
        if (gb) gb[0] = (float)s;
    } else {
        const int t = j / C, c = j - t * C;
        gw[c * 9 + t] = (float)s;
    }
}

inline int log2_ceil(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}
inline int wgrad_blocks(long long rows) { return (int)(rows < 148 * 4 ? (rows < 1 ? 1 : rows) : 148 * 4); }

}  // namespace

cudaError_t dispconv_fwd(const float* xp, const float* w, const float* bias, float* y, int B, int C, int H, int W, cudaStream_t st) {
    const int C4 = C / 4, lg = log2_ceil(C4);
    const long long per_row = (long long)W << lg;
    return launch_pdl(dispconv_fwd_kernel, dim3((unsigned)((per_row + NT - 1) / NT), (unsigned)((B * H + RPB - 1) / RPB)), dim3(NT),
                      (size_t)(9 * C4 * sizeof(float4)), st, (const float4*)xp, w, bias, y, B * H, C4, H, W, lg);
}
cudaError_t dispconv_dgrad(const float* gy, const float* w, float* gxp, int B, int C, int H, int W, cudaStream_t st) {
    const int C4 = C / 4;
    return launch_pdl(dispconv_dgrad_kernel, dim3((unsigned)(((W + 2) * C4 + NT - 1) / NT), (unsigned)((B * (H + 2) + RPB - 1) / RPB)), dim3(NT),
                      (size_t)(9 * C4 * sizeof(float4)), st, gy, w, (float4*)gxp, B * (H + 2), C4, H, W);
}
size_t dispconv_wgrad_workspace_floats(long long P, int C) { return (size_t)(148 * 4) * (9 * C + 4); }
cudaError_t dispconv_wgrad(const float* xp, const float* gy, float* gw, float* gb, float* workspace, int B, int C, int H, int W,
                           cudaStream_t st) {
    const int rows = B * H;
    const int nb = wgrad_blocks(rows), C4 = C / 4, lanes = NT / C4;
    const size_t smem = (size_t)lanes * 9 * C4 * sizeof(float4) + (size_t)lanes * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dispconv_wgrad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaError_t e = launch_pdl(dispconv_wgrad_partial_kernel, dim3(nb), dim3(NT), smem, st, (const float4*)xp, gy, workspace, rows, C4, H, W);
    if (e != cudaSuccess) return e;
    return launch_pdl(dispconv_wgrad_finalize_kernel, dim3((9 * C + 1 + 31) / 32), dim3(1024), 0, st, (const float*)workspace, nb, C, gw, gb);
}

}  // namespace mvf
