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
constexpr int RPB = 4;   // image rows per work item

// Work decomposition (all three kernels).  A work item is RPB image rows x the NT threads of one column block of one image; persistent
// CTAs (two per SM) walk the items.  The 4-channel groups of a pixel sit on neighbouring lanes (tpp = C/4 rounded up to a power of two),
// so a warp's load of one tap is one contiguous run of 32 / tpp pixels.  A thread first issues EVERY load of its item -- the
// (RPB + 2) x 3 window of float4 (fwd, wgrad) or of grad_y scalars (dgrad) -- and only then computes: one exposed memory latency per item
// with 18 independent requests per thread in flight (147 KB per SM).  History at B12 192x640, C = 16 (HBM floor ~16 us each):
//   one thread per pixel walking its 64 B                             126 / 80 / 139 us  (fwd / dgrad / wgrad)
//   lanes = channel groups, 9 loads per row, 8 rows per CTA             76 / 76 /  89 us  (issue- and latency-bound: 3 loads in flight per
//   + 3-row register window (a third of the loads)                       76 / 76 /  89 us   thread, one latency per ROW)
//   every load of an RPB-row item issued before the arithmetic          46 / 40 /  44 us  (2.1 - 2.4 TB/s; a TMA ring would decouple the
//                                                                                          loads from the arithmetic entirely)
__device__ __forceinline__ void stage_weights(float4* ws, const float* __restrict__ w, int C4) {
    for (int i = threadIdx.x; i < 9 * C4; i += NT) {   // [9][C4] tap-major
        const int t = i / C4, c = i - t * C4;
        ws[i] = make_float4(w[(4 * c + 0) * 9 + t], w[(4 * c + 1) * 9 + t], w[(4 * c + 2) * 9 + t], w[(4 * c + 3) * 9 + t]);
    }
}

struct Item {
    int b, yy0, nrows, xb;
};
// item -> (image, first row, rows, column block); `rows_per_image` rows are cut into n_chunk pieces of RPB
__device__ __forceinline__ Item item_of(int item, int n_xblk, int n_chunk, int rows_per_image) {
    Item it;
    it.xb = item % n_xblk;
    const int bc = item / n_xblk;
    it.b = bc / n_chunk;
    it.yy0 = (bc - it.b * n_chunk) * RPB;
    it.nrows = min(RPB, rows_per_image - it.yy0);
    return it;
}

// fwd: a thread owns (pixel column x, channel group c) of RPB output rows
__global__ void __launch_bounds__(NT, 2) dispconv_fwd_kernel(const float4* __restrict__ xp, const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ y, int B, int C4, int H, int W, int tpp_log2, int n_xblk, int n_chunk) {
    pdl_sync();
    extern __shared__ float4 ws[];
    stage_weights(ws, w, C4);
    __syncthreads();
    const int Wp = W + 2;
    const float b0 = bias ? bias[0] : 0.f;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n_items = B * n_chunk * n_xblk;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = item_of(item, n_xblk, n_chunk, H);
        const int t = it.xb * NT + threadIdx.x;
        const int x = t >> tpp_log2, c = t & ((1 << tpp_log2) - 1);
        const bool live = x < W && c < C4;
        const float4* base = xp + (((long long)it.b * (H + 2) + it.yy0) * Wp + x) * C4 + c;
        float4 v[RPB + 2][3];
#pragma unroll
        for (int r = 0; r < RPB + 2; ++r)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) v[r][kw] = (live && r < it.nrows + 2) ? __ldg(base + ((long long)r * Wp + kw) * C4) : z;
        const float4* wq = ws + (live ? c : 0);
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float4 q = wq[(3 * kh + kw) * C4];
                    const float4 u = v[r + kh][kw];
                    a0 = fmaf(u.x, q.x, a0); a1 = fmaf(u.y, q.y, a1); a2 = fmaf(u.z, q.z, a2); a3 = fmaf(u.w, q.w, a3);
                }
            }
            float acc = live ? (a0 + a1) + (a2 + a3) : 0.f;
            for (int o = 1; o < (1 << tpp_log2); o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // fixed butterfly over the pixel's lanes
            if (x < W && c == 0 && r < it.nrows) y[((long long)it.b * H + it.yy0 + r) * W + x] = acc + b0;
        }
    }
}

// dgrad: thread = (padded pixel column px, channel group c) of RPB padded rows; the (RPB + 2) x 3 window of grad_y scalars (zero outside)
__global__ void __launch_bounds__(NT, 2) dispconv_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ w, float4* __restrict__ gxp,
                                                               int B, int C4, int H, int W, int n_xblk, int n_chunk) {
    pdl_sync();
    extern __shared__ float4 ws[];
    stage_weights(ws, w, C4);
    __syncthreads();
    const int Hp = H + 2, Wp = W + 2;
    const int n_items = B * n_chunk * n_xblk;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = item_of(item, n_xblk, n_chunk, Hp);
        const int t = it.xb * NT + threadIdx.x;
        if (t >= Wp * C4) continue;
        const int px = t / C4, c = t - px * C4;
        const float4* wq = ws + c;
        const bool okc[3] = {px < W, px >= 1 && px - 1 < W, px >= 2};   // column px - kw inside [0, W)
        const float* g = gy + (long long)it.b * H * W + px;
        float gv[RPB + 2][3];   // grad_y rows py0 - 2 .. py0 + RPB - 1 at columns px - kw
#pragma unroll
        for (int r = 0; r < RPB + 2; ++r) {
            const int gr = it.yy0 - 2 + r;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) gv[r][kw] = (gr >= 0 && gr < H && r < it.nrows + 2 && okc[kw]) ? __ldg(g + (long long)gr * W - kw) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < RPB; ++j) {
            if (j >= it.nrows) break;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                // output row oy = py - kh: kh = 0 reads grad_y row py (gv[j + 2]), kh = 1 row py - 1, kh = 2 row py - 2
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float4 q = wq[(3 * kh + kw) * C4];
                    const float s = gv[j + 2 - kh][kw];
                    acc.x = fmaf(s, q.x, acc.x); acc.y = fmaf(s, q.y, acc.y); acc.z = fmaf(s, q.z, acc.z); acc.w = fmaf(s, q.w, acc.w);
                }
            }
            gxp[((long long)it.b * Hp + it.yy0 + j) * Wp * C4 + t] = acc;
        }
    }
}

// partial[block][9][C] (+ the bias gradient at [9*C]); a thread owns (pixel column x, channel group c) and keeps its nine float4 sums over
// all the items of its CTA; one block reduction at the end
__global__ void __launch_bounds__(NT, 2) dispconv_wgrad_partial_kernel(const float4* __restrict__ xp, const float* __restrict__ gy,
                                                                       float* __restrict__ partial, int B, int C4, int H, int W, int tpp_log2,
                                                                       int n_xblk, int n_chunk) {
    pdl_sync();
    extern __shared__ float4 red[];   // [lanes][9][C4] + bias column
    const int lanes = NT >> tpp_log2;
    const int l = threadIdx.x >> tpp_log2, c = threadIdx.x & ((1 << tpp_log2) - 1);
    const int Wp = W + 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = z;
    float gsum = 0.f;
    const int n_items = B * n_chunk * n_xblk;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = item_of(item, n_xblk, n_chunk, H);
        const int x = (it.xb * NT + (int)threadIdx.x) >> tpp_log2;
        const bool live = x < W && c < C4;
        const float4* base = xp + (((long long)it.b * (H + 2) + it.yy0) * Wp + x) * C4 + c;
        const float* grow = gy + ((long long)it.b * H + it.yy0) * W + x;
        float4 v[RPB + 2][3];
        float g[RPB];
#pragma unroll
        for (int r = 0; r < RPB + 2; ++r)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) v[r][kw] = (live && r < it.nrows + 2) ? __ldg(base + ((long long)r * Wp + kw) * C4) : z;
#pragma unroll
        for (int r = 0; r < RPB; ++r) g[r] = (live && r < it.nrows) ? __ldg(grow + (long long)r * W) : 0.f;
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
            if (c == 0) gsum += g[r];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 u = v[r + kh][kw];
                    float4& a = acc[kh * 3 + kw];
                    a.x = fmaf(g[r], u.x, a.x); a.y = fmaf(g[r], u.y, a.y); a.z = fmaf(g[r], u.z, a.z); a.w = fmaf(g[r], u.w, a.w);
                }
        }
    }
    if (c < C4) {
#pragma unroll
        for (int t = 0; t < 9; ++t) red[(l * 9 + t) * C4 + c] = acc[t];
    }
    float* gred = reinterpret_cast<float*>(red + (size_t)lanes * 9 * C4);
    if (c == 0) gred[l] = gsum;
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
constexpr int MAX_CTAS = 148 * 4;   // rows of the weight-gradient workspace
inline int persistent_ctas(long long n_items, int per_sm = 2) {
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    const long long g = (long long)per_sm * n_sm < MAX_CTAS ? (long long)per_sm * n_sm : MAX_CTAS;   // resident CTAs per SM (registers)
    return (int)(n_items < g ? (n_items < 1 ? 1 : n_items) : g);
}

}  // namespace

cudaError_t dispconv_fwd(const float* xp, const float* w, const float* bias, float* y, int B, int C, int H, int W, cudaStream_t st) {
    const int C4 = C / 4, lg = log2_ceil(C4);
    const int n_xblk = (int)((((long long)W << lg) + NT - 1) / NT), n_chunk = (H + RPB - 1) / RPB;
    return launch_pdl(dispconv_fwd_kernel, dim3((unsigned)persistent_ctas((long long)B * n_chunk * n_xblk)), dim3(NT), (size_t)(9 * C4 * sizeof(float4)), st,
                      (const float4*)xp, w, bias, y, B, C4, H, W, lg, n_xblk, n_chunk);
}
cudaError_t dispconv_dgrad(const float* gy, const float* w, float* gxp, int B, int C, int H, int W, cudaStream_t st) {
    const int C4 = C / 4;
    const int n_xblk = ((W + 2) * C4 + NT - 1) / NT, n_chunk = (H + 2 + RPB - 1) / RPB;
    // 79 registers: three CTAs per SM
    return launch_pdl(dispconv_dgrad_kernel, dim3((unsigned)persistent_ctas((long long)B * n_chunk * n_xblk, 3)), dim3(NT), (size_t)(9 * C4 * sizeof(float4)), st,
                      gy, w, (float4*)gxp, B, C4, H, W, n_xblk, n_chunk);
}
size_t dispconv_wgrad_workspace_floats(long long P, int C) { return (size_t)MAX_CTAS * (9 * C + 4); }
cudaError_t dispconv_wgrad(const float* xp, const float* gy, float* gw, float* gb, float* workspace, int B, int C, int H, int W,
                           cudaStream_t st) {
    const int C4 = C / 4, lg = log2_ceil(C4), lanes = NT >> lg;
    const int n_xblk = (int)((((long long)W << lg) + NT - 1) / NT), n_chunk = (H + RPB - 1) / RPB;
    const int nb = persistent_ctas((long long)B * n_chunk * n_xblk);
    const size_t smem = (size_t)lanes * 9 * C4 * sizeof(float4) + (size_t)lanes * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dispconv_wgrad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaError_t e = launch_pdl(dispconv_wgrad_partial_kernel, dim3(nb), dim3(NT), smem, st, (const float4*)xp, gy, workspace, B, C4, H, W, lg, n_xblk,
                               n_chunk);
    if (e != cudaSuccess) return e;
    return launch_pdl(dispconv_wgrad_finalize_kernel, dim3((9 * C + 1 + 31) / 32), dim3(1024), 0, st, (const float*)workspace, nb, C, gw, gb);
}

}  // namespace mvf
