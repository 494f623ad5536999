// Gradient clipping + AdamW over ONE flat fp32 arena (parameters, gradients, both moments): SURVEY.md 8(f) item 1.
// Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step (train.py:661-666), which walk ~200 tensors with a
// dozen multi-tensor launches, by two passes over the arena:
//   1. sum of squares of the gradient (per-CTA partials, fixed-order final sum -> deterministic)
//   2. clip coefficient min(1, max_norm / (norm + 1e-6)) and the AdamW update, step counter kept on the device so the
//      whole thing can be recorded into a CUDA graph.
// HBM-bound: pass 1 reads 4 B/param, pass 2 reads 16 and writes 12 B/param.
#include "optim.cuh"

#include <stdint.h>

namespace mvf {
namespace {

constexpr int NT = 256;
constexpr int NB_NORM = 296;  // 2 CTAs per SM

// Sum of squares of the (scaled) gradient.  The first n_dup4 float4 groups belong to parameters the reference lists TWICE
// in its optimizer (train.py:198-200: `encoder` and `encoder_mf` alias one module under shared_encoder / shared_all), so
// clip_grad_norm_ counts their norm twice: they are added twice here.
__global__ void __launch_bounds__(NT) sumsq_partial_kernel(const float4* __restrict__ g, long long n4, long long n_dup4,
                                                           const float* __restrict__ tail, int ntail, double* __restrict__ partial) {
    double acc = 0.0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        const float4 v = __ldg(g + i);
        if (i < n_dup4) { d0 = fmaf(v.x, v.x, d0); d1 = fmaf(v.y, v.y, d1); d2 = fmaf(v.z, v.z, d2); d3 = fmaf(v.w, v.w, d3); }
        else { a0 = fmaf(v.x, v.x, a0); a1 = fmaf(v.y, v.y, a1); a2 = fmaf(v.z, v.z, a2); a3 = fmaf(v.w, v.w, a3); }
    }
    acc = (double)a0 + (double)a1 + (double)a2 + (double)a3 + 2.0 * ((double)d0 + (double)d1 + (double)d2 + (double)d3);
    if (blockIdx.x == 0 && threadIdx.x < ntail) acc += (double)tail[threadIdx.x] * (double)tail[threadIdx.x];
    __shared__ double red[NT];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) {  // fixed tree
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// state (device, 4 floats): [0] step count, [1] last gradient norm (for logging), [2] learning rate, [3] gradient scale
// (1 / world size after a sum-all-reduce, else 1).  Learning rate and scale live on the device so that a recorded CUDA
// graph follows the scheduler (MultiStepLR / cosine, train.py:239-242) without being re-captured.
__global__ void __launch_bounds__(NT) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, long long n_dup4,
                                                   const unsigned char* __restrict__ skip, const double* __restrict__ partial, int nb,
                                                   const float* __restrict__ state, float beta1, float beta2, float eps,
                                                   float wd, float max_norm) {
    __shared__ float s_coef, s_lr, s_bc1[3], s_bc2s[3];
    if (threadIdx.x == 0) {
        double ss = 0.0;
        for (int b = 0; b < nb; ++b) ss += partial[b];
        const float gscale = state[3];
        const float norm = (float)sqrt(ss) * gscale;
        float coef = 1.0f;
        if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (norm + 1e-6f));  // clip_grad_norm_
        s_coef = coef * gscale;
        s_lr = state[2];
        // ordinary parameters are at step t = state[0] + 1; duplicated ones take two updates per iteration, their own
        // counter running at 2t - 1 and 2t (torch keeps one state per tensor and walks the parameter list)
        const float t = state[0] + 1.0f;
        const float ts[3] = {t, 2.0f * t - 1.0f, 2.0f * t};
        for (int k = 0; k < 3; ++k) {
            s_bc1[k] = 1.0f - powf(beta1, ts[k]);
            s_bc2s[k] = sqrtf(1.0f - powf(beta2, ts[k]));
        }
    }
    __syncthreads();
    const float coef = s_coef, lr = s_lr;
    // a duplicated gradient is scaled by the clip coefficient twice (clip_grad_norm_ multiplies every list entry)
    const float clip_only = (state[3] != 0.f) ? coef / state[3] : 1.0f;
    const long long n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    auto upd = [&](float& pp, float gg, float& mm, float& vv, int k) {
        pp *= (1.0f - lr * wd);
        mm = beta1 * mm + (1.0f - beta1) * gg;   // lerp form of torch: m + (g - m)(1 - beta1)
        vv = beta2 * vv + (1.0f - beta2) * gg * gg;
        const float denom = sqrtf(vv) / s_bc2s[k] + eps;
        pp -= (lr / s_bc1[k]) * (mm / denom);
    };
    auto upd_any = [&](float& pp, float gg, float& mm, float& vv, bool dup) {
        if (!dup) {
            upd(pp, gg * coef, mm, vv, 0);
        } else {
            gg *= coef * clip_only;
            upd(pp, gg, mm, vv, 1);
            upd(pp, gg, mm, vv, 2);
        }
    };
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        if (skip != nullptr && skip[i]) continue;   // parameter without a gradient this step: untouched, as torch leaves it
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        const bool dup = i < n_dup4;
        upd_any(pp.x, gg.x, mm.x, vv.x, dup); upd_any(pp.y, gg.y, mm.y, vv.y, dup);
        upd_any(pp.z, gg.z, mm.z, vv.z, dup); upd_any(pp.w, gg.w, mm.w, vv.w, dup);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    if (blockIdx.x == 0) {
        const long long i = n4 * 4 + threadIdx.x;
        if (i < n) upd_any(p[i], g[i], m[i], v[i], false);
    }
}

// bumps the step counter AFTER every CTA of adamw_kernel has read it (separate tiny launch on the same stream)
__global__ void adamw_tick_kernel(float* state, const double* partial, int nb) {
    double ss = 0.0;
    for (int b = 0; b < nb; ++b) ss += partial[b];
    state[0] += 1.0f;
    state[1] = (float)sqrt(ss) * state[3];
}

// Multi-tensor gather: autograd leaves one gradient tensor per parameter; this copies up to GATHER_MAX of them per launch
// into their slices of the flat arena (blockIdx.y = tensor, blockIdx.x strides over it), so the parameters' .grad can
// stay None between steps -- autograd then adopts each gradient instead of launching one `grad += new` per parameter.
constexpr int GATHER_MAX = 128;
struct GatherTable {
    const float* src[GATHER_MAX];
    long long off[GATHER_MAX];
    long long n[GATHER_MAX];
};

__global__ void __launch_bounds__(NT) gather_grads_kernel(const __grid_constant__ GatherTable t, float* __restrict__ G) {
    const float* __restrict__ src = t.src[blockIdx.y];
    float* __restrict__ dst = G + t.off[blockIdx.y];
    const long long n = t.n[blockIdx.y];
    if (src == nullptr) {  // parameter without a gradient this step
        for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) dst[i] = 0.f;
        return;
    }
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {  // arena slices are 16-byte aligned by construction
        const long long n4 = n / 4;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) d4[i] = __ldg(s4 + i);
        if (blockIdx.x == 0) {
            const long long i = n4 * 4 + threadIdx.x;
            if (i < n) dst[i] = src[i];
        }
    } else {
        for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) dst[i] = src[i];
    }
}

}  // namespace

cudaError_t gather_grads(float* G, const void* const* srcs, const long long* offsets, const long long* sizes, int n_tensors,
                         cudaStream_t st) {
    for (int base = 0; base < n_tensors; base += GATHER_MAX) {
        GatherTable t;
        const int cnt = n_tensors - base < GATHER_MAX ? n_tensors - base : GATHER_MAX;
        long long nmax = 1;
        for (int i = 0; i < GATHER_MAX; ++i) {
            const bool ok = i < cnt;
            t.src[i] = ok ? reinterpret_cast<const float*>(srcs[base + i]) : nullptr;
            t.off[i] = ok ? offsets[base + i] : 0;
            t.n[i] = ok ? sizes[base + i] : 0;
            if (ok && sizes[base + i] > nmax) nmax = sizes[base + i];
        }
        long long bx = (nmax / 4 + NT * 8 - 1) / (NT * 8);  // ~8 float4 per thread on the largest tensor
        if (bx < 1) bx = 1;
        if (bx > 64) bx = 64;
        gather_grads_kernel<<<dim3((unsigned)bx, (unsigned)cnt), NT, 0, st>>>(t, G);
    }
    return cudaGetLastError();
}

size_t adamw_workspace_bytes() { return NB_NORM * sizeof(double); }

cudaError_t adamw_step(float* p, const float* g, float* m, float* v, long long n, long long n_dup, const unsigned char* skip,
                       float* state, void* workspace, float beta1, float beta2, float eps, float wd, float max_norm,
                       cudaStream_t st) {
    double* partial = reinterpret_cast<double*>(workspace);
    const long long n4 = n / 4;
    sumsq_partial_kernel<<<NB_NORM, NT, 0, st>>>(reinterpret_cast<const float4*>(g), n4, n_dup / 4, g + n4 * 4, (int)(n - n4 * 4),
                                                 partial);
    long long nb = (n4 + NT - 1) / NT;
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    adamw_kernel<<<(int)nb, NT, 0, st>>>(p, g, m, v, n, n_dup / 4, skip, partial, NB_NORM, state, beta1, beta2, eps, wd, max_norm);
    adamw_tick_kernel<<<1, 1, 0, st>>>(state, partial, NB_NORM);
    return cudaGetLastError();
}

}  // namespace mvf
