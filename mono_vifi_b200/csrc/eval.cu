// Evaluation path (train.py:419-483 `test_kitti`, evaluate_depth.py:134-193), inference only:
//   bn_eval      BatchNorm2d in eval mode (running statistics) + residual add + ReLU, channels-last -- the inference form of the
//                bn -> (+= identity) -> relu tails of the encoders;
//   depth_eval   per image: bilinear resize of the predicted (scaled) disparity to the ground-truth size (align_corners=False),
//                depth = 1 / disparity, validity mask (min < gt < max and the Eigen crop, or gt > 0), median scaling
//                (torch.median: the lower median) or the fixed stereo factor, clamp, and the seven error metrics of
//                compute_depth_errors (layers.py:293-311).  Three launches per image, no host synchronisation: the medians come
//                from an exact 4-pass radix select over the valid pixels, the sums from one CTA in a fixed order (reproducible).
#include "eval.cuh"

#include <cstdint>

#include "pdl.cuh"

namespace mvf {
namespace {

constexpr int NT = 256;

__global__ void bn_eval_kernel(const float4* __restrict__ x, const float4* __restrict__ identity, float4* __restrict__ y,
                               const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                               const float* __restrict__ beta, long long total4, int C4, float eps, int relu) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total4; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        const float4 m = *reinterpret_cast<const float4*>(mean + 4 * c), vv = *reinterpret_cast<const float4*>(var + 4 * c);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + 4 * c), be = *reinterpret_cast<const float4*>(beta + 4 * c);
        const float4 v = __ldg(x + i);
        float4 o;
        o.x = fmaf((v.x - m.x) * rsqrtf(vv.x + eps), ga.x, be.x);
        o.y = fmaf((v.y - m.y) * rsqrtf(vv.y + eps), ga.y, be.y);
        o.z = fmaf((v.z - m.z) * rsqrtf(vv.z + eps), ga.z, be.z);
        o.w = fmaf((v.w - m.w) * rsqrtf(vv.w + eps), ga.w, be.w);
        if (identity) {
            const float4 d = __ldg(identity + i);
            o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        y[i] = o;
    }
}

// ---- depth metrics -----------------------------------------------------------------------------------------------------------
__global__ void depth_prepare_kernel(const float* __restrict__ disp, const float* __restrict__ gt, float* __restrict__ pred,
                                     unsigned char* __restrict__ mask, int h, int w, int Hg, int Wg, float sh, float sw, float min_d,
                                     float max_d, int y0, int y1, int x0, int x1, int eigen) {
    const long long total = (long long)Hg * Wg;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int ox = (int)(i % Wg), oy = (int)(i / Wg);
        // F.interpolate(mode="bilinear", align_corners=False): torch's source index and weights
        const float sy = fmaxf(fmaf(sh, (float)oy + 0.5f, -0.5f), 0.f), sx = fmaxf(fmaf(sw, (float)ox + 0.5f, -0.5f), 0.f);
        const int iy0 = min((int)sy, h - 1), ix0 = min((int)sx, w - 1);
        const int iy1 = iy0 + (iy0 < h - 1 ? 1 : 0), ix1 = ix0 + (ix0 < w - 1 ? 1 : 0);
        const float ly = sy - (float)iy0, lx = sx - (float)ix0;
        const float top = (1.f - lx) * __ldg(disp + iy0 * w + ix0) + lx * __ldg(disp + iy0 * w + ix1);
        const float bot = (1.f - lx) * __ldg(disp + iy1 * w + ix0) + lx * __ldg(disp + iy1 * w + ix1);
        pred[i] = 1.0f / ((1.f - ly) * top + ly * bot);
        const float g = __ldg(gt + i);
        const bool ok = eigen ? (g > min_d && g < max_d && oy >= y0 && oy < y1 && ox >= x0 && ox < x1) : (g > 0.f);
        mask[i] = ok ? 1 : 0;
    }
}

// order-preserving key of a float (all values of interest are positive, but the map is the general one)
__device__ __forceinline__ uint32_t fkey(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// block b: exact lower median (torch.median) of the valid entries of array b (0: gt, 1: pred); out[b] = median, out[2] = count
__global__ void __launch_bounds__(1024) masked_median_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                             const unsigned char* __restrict__ mask, long long n, float* __restrict__ out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_k, s_count;
    const float* v = blockIdx.x == 0 ? gt : pred;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    unsigned int cnt = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) cnt += mask[i];
    atomicAdd(&s_count, cnt);
    __syncthreads();
    const unsigned int count = s_count;
    if (count == 0) {
        if (threadIdx.x == 0) {
            out[blockIdx.x] = __uint_as_float(0x7fc00000u);
            out[2] = 0.f;
        }
        return;
    }
    if (threadIdx.x == 0) {
        s_prefix = 0;
        s_k = (count - 1) / 2;   // 0-based rank of the lower median
    }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int j = threadIdx.x; j < 256; j += blockDim.x) hist[j] = 0;
        __syncthreads();
        const unsigned int prefix = s_prefix;
        const unsigned int hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            if (!mask[i]) continue;
            const uint32_t k = fkey(v[i]);
            if ((k & hi_mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int k = s_k, b = 0;
            while (b < 255 && k >= hist[b]) {
                k -= hist[b];
                ++b;
            }
            s_k = k;
            s_prefix = prefix | (b << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[blockIdx.x] = fkey_inv(s_prefix);
        out[2] = (float)count;
    }
}

// one CTA: metrics[0..6] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3; metrics[7] = the scale ratio applied
__global__ void __launch_bounds__(1024) depth_errors_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                            const unsigned char* __restrict__ mask, long long n,
                                                            const float* __restrict__ med, float stereo_scale, float min_d, float max_d,
                                                            float* __restrict__ metrics) {
    __shared__ double red[7][32];
    const float ratio = stereo_scale > 0.f ? stereo_scale : med[0] / med[1];
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        if (!mask[i]) continue;
        const float g = gt[i];
        const float p = fminf(fmaxf(pred[i] * ratio, min_d), max_d);
        const float th = fmaxf(g / p, p / g), d = g - p, dl = logf(g) - logf(p);
        s[0] += (double)(fabsf(d) / g);
        s[1] += (double)(d * d / g);
        s[2] += (double)(d * d);
        s[3] += (double)(dl * dl);
        s[4] += th < 1.25f ? 1.0 : 0.0;
        s[5] += th < 1.25f * 1.25f ? 1.0 : 0.0;
        s[6] += th < 1.25f * 1.25f * 1.25f ? 1.0 : 0.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        double v = s[q];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0.0;
        for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) v += red[threadIdx.x][wi];
        const double cnt = (double)med[2];
        v = cnt > 0 ? v / cnt : 0.0;
        if (threadIdx.x == 2 || threadIdx.x == 3) v = sqrt(v);
        metrics[threadIdx.x] = (float)v;
    }
    if (threadIdx.x == 7) metrics[7] = ratio;
}

}  // namespace

cudaError_t bn_eval_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, long long P, int C, float eps, int relu, cudaStream_t st) {
    const long long total4 = P * (C / 4);
    long long g = (total4 + NT - 1) / NT;
    if (g > 148 * 16) g = 148 * 16;
    return launch_pdl(bn_eval_kernel, dim3((unsigned)(g < 1 ? 1 : g)), dim3(NT), 0, st, (const float4*)x, (const float4*)identity, (float4*)y,
                      running_mean, running_var, gamma, beta, total4, C / 4, eps, relu);
}

size_t depth_eval_workspace_bytes(int Hg, int Wg) { return (size_t)Hg * Wg * 5 + 64; }

cudaError_t depth_eval(const float* disp, int h, int w, const float* gt, int Hg, int Wg, float min_d, float max_d, int eigen_crop,
                       float stereo_scale, void* workspace, float* metrics8, cudaStream_t st) {
    const long long n = (long long)Hg * Wg;
    float* pred = reinterpret_cast<float*>(workspace);
    float* med = pred + n;                                   // 3 floats (+ padding)
    unsigned char* mask = reinterpret_cast<unsigned char*>(med + 4);
    // the Eigen crop of train.py:452-456 (python int() of the products)
    const int y0 = (int)(0.40810811 * Hg), y1 = (int)(0.99189189 * Hg), x0 = (int)(0.03594771 * Wg), x1 = (int)(0.96405229 * Wg);
    const float sh = (float)h / (float)Hg, sw = (float)w / (float)Wg;
    long long g = (n + NT - 1) / NT;
    if (g > 148 * 8) g = 148 * 8;
    depth_prepare_kernel<<<(unsigned)g, NT, 0, st>>>(disp, gt, pred, mask, h, w, Hg, Wg, sh, sw, min_d, max_d, y0, y1, x0, x1, eigen_crop);
    masked_median_kernel<<<2, 1024, 0, st>>>(gt, pred, mask, n, med);
    depth_errors_kernel<<<1, 1024, 0, st>>>(gt, pred, mask, n, med, stereo_scale, min_d, max_d, metrics8);
    return cudaGetLastError();
}

}  // namespace mvf
