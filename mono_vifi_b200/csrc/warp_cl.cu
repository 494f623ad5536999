// Feature warp ("F2"), bilinear resize, PReLU tail and pose-matrix kernels: the HBM-bound glue of the multi-frame
// branch that the reference runs as F.grid_sample / F.interpolate / nn.PReLU / ~40 tiny ATen launches per pose matrix.
//
//   flow_warp      IFRNet.warp (IFRNet.py:7-15): backward warp of an image or a feature map by a pixel-unit flow, bilinear, border
//                  padding, align_corners=True; used on the VFI pyramids, the final frame synthesis and by
//                  FusionModule.warp_features (fusion_module.py:78-90).  Backward to the warped tensor (the flows come from the
//                  frozen VFI network): the reference scatters with float atomicAdd (order-dependent); here the scatter accumulates
//                  64-bit fixed point (scale = a power of two chosen from max|grad|, so the integer sums are order-independent),
//                  then one pass converts back to fp32 -- bitwise deterministic.
//   resize_bilinear  F.interpolate(mode="bilinear") with align_corners True (hrnet_encoder.py:275-280) or False (IFRNet.py:118,
//                  383-423, fusion_module.py:68-99, LiteMono.py:495,502 through layers.upsample), any size ratio, optional
//                  per-channel multiplier (flow rescaling).  Backward in gather form: every input pixel sums the output pixels
//                  whose footprint contains it, with the forward's own index/weight function -- exact adjoint, no atomics.
//   prelu_cl       nn.PReLU(C) (+ the residual add of IFRNet's ResBlock, IFRNet.py:140-150), channels-last.
//   pose_matrix    transformation_from_parameters (layers.py:28-103): Rodrigues rotation, translation, (inverse) product.
// Layouts: 0 = dense NCHW, one element per thread; 1 = dense channels-last, one float4 (C % 4 == 0) or float2 (C % 2 == 0) per thread.
#include "warp_cl.cuh"

#include <cstdint>

#include "pdl.cuh"

namespace mvf {
namespace {

constexpr int NT = 256;

inline int grid_for(long long total) {
    long long g = (total + NT - 1) / NT;
    const long long cap = 148LL * 16;   // a few waves of the 148 SMs; grid-stride beyond
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---- source index of torch's upsample_bilinear2d (area_pixel_compute_source_index, cubic = false) ----------------------------
struct Tap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Tap tap_of(float scale, int dst, int in_size, int align) {
    float s = align ? scale * (float)dst : fmaxf(fmaf(scale, (float)dst + 0.5f, -0.5f), 0.f);
    Tap t;
    t.i0 = min((int)s, in_size - 1);
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = s - (float)t.i0;
    t.l0 = 1.f - t.l1;
    return t;
}

// channels-last vectors: V = 4 (float4, C % 4 == 0) or V = 2 (float2: HRNet's 18-channel branch)
template <int V>
struct Vec {
    float v[V];
};
template <int V>
__device__ __forceinline__ Vec<V> vld(const float* p) {
    Vec<V> r;
    if constexpr (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    }
    return r;
}
template <int V>
__device__ __forceinline__ void vst(float* p, const Vec<V>& a) {
    if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
    else *reinterpret_cast<float2*>(p) = make_float2(a.v[0], a.v[1]);
}
template <int V>
__device__ __forceinline__ Vec<V> vzero() {
    Vec<V> r;
#pragma unroll
    for (int j = 0; j < V; ++j) r.v[j] = 0.f;
    return r;
}
template <int V>
__device__ __forceinline__ Vec<V> vfma(float w, const Vec<V>& x, const Vec<V>& a) {
    Vec<V> r;
#pragma unroll
    for (int j = 0; j < V; ++j) r.v[j] = fmaf(w, x.v[j], a.v[j]);
    return r;
}
template <int V>
__device__ __forceinline__ Vec<V> vmul(float w, const Vec<V>& x) {
    Vec<V> r;
#pragma unroll
    for (int j = 0; j < V; ++j) r.v[j] = w * x.v[j];
    return r;
}
template <int V>
__device__ __forceinline__ Vec<V> vadd(const Vec<V>& a, const Vec<V>& b) {
    Vec<V> r;
#pragma unroll
    for (int j = 0; j < V; ++j) r.v[j] = a.v[j] + b.v[j];
    return r;
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ---- resize, forward ---------------------------------------------------------------------------------------------------------
template <int V>
__global__ void resize_fwd_cl_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C4, int Hi, int Wi, int Ho, int Wo,
                                     float sh, float sw, int align) {
    pdl_sync();
    const long long total = (long long)B * Ho * Wo * C4;   // C4 = C / V vectors per pixel
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int ox = (int)(r % Wo);
        r /= Wo;
        const int oy = (int)(r % Ho), b = (int)(r / Ho);
        const Tap ty = tap_of(sh, oy, Hi, align), tx = tap_of(sw, ox, Wi, align);
        const float* p = x + ((long long)b * Hi * Wi * C4 + c) * V;
        const long long cs = (long long)C4 * V;
        const Vec<V> v00 = vld<V>(p + ((long long)ty.i0 * Wi + tx.i0) * cs), v01 = vld<V>(p + ((long long)ty.i0 * Wi + tx.i1) * cs);
        const Vec<V> v10 = vld<V>(p + ((long long)ty.i1 * Wi + tx.i0) * cs), v11 = vld<V>(p + ((long long)ty.i1 * Wi + tx.i1) * cs);
        const Vec<V> top = vadd(vmul(tx.l0, v00), vmul(tx.l1, v01)), bot = vadd(vmul(tx.l0, v10), vmul(tx.l1, v11));
        vst<V>(y + i * V, vadd(vmul(ty.l0, top), vmul(ty.l1, bot)));
    }
}

__global__ void resize_fwd_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int Hi, int Wi, int Ho, int Wo,
                                       float sh, float sw, int align, float mul_even, float mul_odd) {
    pdl_sync();
    const long long total = (long long)B * C * Ho * Wo;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int ox = (int)(i % Wo);
        long long r = i / Wo;
        const int oy = (int)(r % Ho);
        r /= Ho;   // r = b * C + c
        const int c = (int)(r % C);
        const Tap ty = tap_of(sh, oy, Hi, align), tx = tap_of(sw, ox, Wi, align);
        const float* p = x + r * Hi * Wi;
        const float top = tx.l0 * __ldg(p + (long long)ty.i0 * Wi + tx.i0) + tx.l1 * __ldg(p + (long long)ty.i0 * Wi + tx.i1);
        const float bot = tx.l0 * __ldg(p + (long long)ty.i1 * Wi + tx.i0) + tx.l1 * __ldg(p + (long long)ty.i1 * Wi + tx.i1);
        y[i] = (ty.l0 * top + ty.l1 * bot) * ((c & 1) ? mul_odd : mul_even);
    }
}

// ---- resize, backward (gather form) ------------------------------------------------------------------------------------------
// output indices whose footprint can contain input index `i`: source coordinate in (i - 1, i + 1); one extra on each side covers the
// rounding of the inverse map, and every candidate is re-tested with the forward's tap function
__device__ __forceinline__ void out_range(float scale, int i, int out_size, int align, int& lo, int& hi) {
    if (scale <= 0.f) {   // align_corners with a single output element
        lo = 0;
        hi = out_size - 1;
        return;
    }
    const float inv = 1.f / scale;
    float a = align ? ((float)i - 1.f) * inv : ((float)i - 0.5f) * inv - 0.5f;
    float b = align ? ((float)i + 1.f) * inv : ((float)i + 1.5f) * inv - 0.5f;
    lo = max(0, (int)floorf(a) - 1);
    hi = min(out_size - 1, (int)ceilf(b) + 1);
}
__device__ __forceinline__ float tap_weight(const Tap& t, int i) { return (t.i0 == i ? t.l0 : 0.f) + (t.i1 == i ? t.l1 : 0.f); }

template <int V>
__global__ void resize_bwd_cl_kernel(const float* __restrict__ gy, float* __restrict__ gx, int B, int C4, int Hi, int Wi, int Ho, int Wo,
                                     float sh, float sw, int align) {
    pdl_sync();
    const long long total = (long long)B * Hi * Wi * C4;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int ix = (int)(r % Wi);
        r /= Wi;
        const int iy = (int)(r % Hi), b = (int)(r / Hi);
        int ylo, yhi, xlo, xhi;
        out_range(sh, iy, Ho, align, ylo, yhi);
        out_range(sw, ix, Wo, align, xlo, xhi);
        const float* p = gy + ((long long)b * Ho * Wo * C4 + c) * V;
        const long long cs = (long long)C4 * V;
        Vec<V> acc = vzero<V>();
        for (int oy = ylo; oy <= yhi; ++oy) {
            const float wy = tap_weight(tap_of(sh, oy, Hi, align), iy);
            if (wy == 0.f) continue;
            Vec<V> row = vzero<V>();
            for (int ox = xlo; ox <= xhi; ++ox) {
                const float wx = tap_weight(tap_of(sw, ox, Wi, align), ix);
                if (wx != 0.f) row = vfma(wx, vld<V>(p + ((long long)oy * Wo + ox) * cs), row);
            }
            acc = vfma(wy, row, acc);
        }
        vst<V>(gx + i * V, acc);
    }
}

__global__ void resize_bwd_nchw_kernel(const float* __restrict__ gy, float* __restrict__ gx, int B, int C, int Hi, int Wi, int Ho, int Wo,
                                       float sh, float sw, int align, float mul_even, float mul_odd) {
    pdl_sync();
    const long long total = (long long)B * C * Hi * Wi;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int ix = (int)(i % Wi);
        long long r = i / Wi;
        const int iy = (int)(r % Hi);
        r /= Hi;
        const int c = (int)(r % C);
        int ylo, yhi, xlo, xhi;
        out_range(sh, iy, Ho, align, ylo, yhi);
        out_range(sw, ix, Wo, align, xlo, xhi);
        const float* p = gy + r * Ho * Wo;
        float acc = 0.f;
        for (int oy = ylo; oy <= yhi; ++oy) {
            const float wy = tap_weight(tap_of(sh, oy, Hi, align), iy);
            if (wy == 0.f) continue;
            float row = 0.f;
            for (int ox = xlo; ox <= xhi; ++ox) {
                const float wx = tap_weight(tap_of(sw, ox, Wi, align), ix);
                if (wx != 0.f) row = fmaf(wx, __ldg(p + (long long)oy * Wo + ox), row);
            }
            acc = fmaf(wy, row, acc);
        }
        gx[i] = acc * ((c & 1) ? mul_odd : mul_even);
    }
}

// ---- flow warp ---------------------------------------------------------------------------------------------------------------
// sampling position of output pixel (y, x): torch.linspace(-1, 1, n)[i] + flow / ((n - 1) / 2), un-normalised and clamped the way
// grid_sample(align_corners=True, padding_mode="border") does it; separately rounded fp32 operations, as the reference's op chain
struct Corner {
    int x0, y0;
    float wnw, wne, wsw, wse;
    bool xin, yin;   // x0 + 1 / y0 + 1 inside the image (x0, y0 always are after the border clamp)
};
__device__ __forceinline__ float linspace_pm1(int i, int n) {
    const float step = 2.0f / (float)(n - 1);
    return i < n / 2 ? __fadd_rn(-1.0f, __fmul_rn(step, (float)i)) : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
}
__device__ __forceinline__ Corner corner_of(float fx, float fy, int x, int y, int W, int H) {
    const float gx = __fadd_rn(linspace_pm1(x, W), __fdiv_rn(fx, ((float)W - 1.0f) / 2.0f));
    const float gy = __fadd_rn(linspace_pm1(y, H), __fdiv_rn(fy, ((float)H - 1.0f) / 2.0f));
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(W - 1));
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(H - 1));
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    const float x0f = floorf(ix), y0f = floorf(iy);
    Corner c;
    c.x0 = (int)x0f;
    c.y0 = (int)y0f;
    const float ex = __fsub_rn(x0f + 1.f, ix), ey = __fsub_rn(y0f + 1.f, iy), dx = __fsub_rn(ix, x0f), dy = __fsub_rn(iy, y0f);
    c.wnw = __fmul_rn(ex, ey);
    c.wne = __fmul_rn(dx, ey);
    c.wsw = __fmul_rn(ex, dy);
    c.wse = __fmul_rn(dx, dy);
    c.xin = c.x0 + 1 < W;
    c.yin = c.y0 + 1 < H;
    return c;
}

template <int V>
__global__ void flow_warp_fwd_cl_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ y, int B, int C4,
                                        int H, int W) {
    pdl_sync();
    const long long total = (long long)B * H * W * C4;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int px = (int)(r % W);
        r /= W;
        const int py = (int)(r % H), b = (int)(r / H);
        const float* f = flow + (long long)b * 2 * H * W + (long long)py * W + px;
        const Corner k = corner_of(__ldg(f), __ldg(f + (long long)H * W), px, py, W, H);
        const float* p = x + (((long long)b * H * W) * C4 + c) * V;
        const long long cs = (long long)C4 * V;
        Vec<V> acc = vmul(k.wnw, vld<V>(p + ((long long)k.y0 * W + k.x0) * cs));
        if (k.xin) acc = vfma(k.wne, vld<V>(p + ((long long)k.y0 * W + k.x0 + 1) * cs), acc);
        if (k.yin) acc = vfma(k.wsw, vld<V>(p + ((long long)(k.y0 + 1) * W + k.x0) * cs), acc);
        if (k.xin && k.yin) acc = vfma(k.wse, vld<V>(p + ((long long)(k.y0 + 1) * W + k.x0 + 1) * cs), acc);
        vst<V>(y + i * V, acc);
    }
}

__global__ void flow_warp_fwd_nchw_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ y, int B, int C,
                                          int H, int W) {
    pdl_sync();
    const long long total = (long long)B * H * W;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int px = (int)(i % W);
        long long r = i / W;
        const int py = (int)(r % H), b = (int)(r / H);
        const float* f = flow + (long long)b * 2 * H * W + (long long)py * W + px;
        const Corner k = corner_of(__ldg(f), __ldg(f + (long long)H * W), px, py, W, H);
        for (int c = 0; c < C; ++c) {
            const float* p = x + ((long long)b * C + c) * H * W;
            float acc = k.wnw * __ldg(p + (long long)k.y0 * W + k.x0);
            if (k.xin) acc = fmaf(k.wne, __ldg(p + (long long)k.y0 * W + k.x0 + 1), acc);
            if (k.yin) acc = fmaf(k.wsw, __ldg(p + (long long)(k.y0 + 1) * W + k.x0), acc);
            if (k.xin && k.yin) acc = fmaf(k.wse, __ldg(p + (long long)(k.y0 + 1) * W + k.x0 + 1), acc);
            y[((long long)b * C + c) * H * W + (long long)py * W + px] = acc;
        }
    }
}

// backward: workspace = [ unsigned absmax bits | pad to 16 B | int64 accumulators B*H*W*C ]
__global__ void absmax_kernel(const float2* __restrict__ g, long long n2, unsigned* __restrict__ out) {
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n2; i += (long long)gridDim.x * NT) {
        const float2 v = __ldg(g + i);
        m = fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y)));
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bit patterns
}
// power-of-two scale that puts max|g| at ~2^40: 2^22 contributions of that size still fit in 63 bits; scaling by it is exact
__device__ __forceinline__ float fixed_scale(unsigned absmax_bits) {
    const int e = (int)((absmax_bits >> 23) & 0xff) - 127;   // floor(log2(max|g|)); 0 / denormal -> -127
    int s = 40 - e;
    s = s > 126 ? 126 : (s < -126 ? -126 : s);
    return __uint_as_float((unsigned)(s + 127) << 23);
}
__device__ __forceinline__ void fx_add(long long* p, float v, float scale) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * scale));
}
template <int V>
__global__ void flow_warp_scatter_kernel(const float* __restrict__ gy, const float* __restrict__ flow, long long* __restrict__ acc,
                                         const unsigned* __restrict__ absmax, int B, int C4, int H, int W) {
    const float scale = fixed_scale(*absmax);
    const long long total = (long long)B * H * W * C4;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int px = (int)(r % W);
        r /= W;
        const int py = (int)(r % H), b = (int)(r / H);
        const float* f = flow + (long long)b * 2 * H * W + (long long)py * W + px;
        const Corner k = corner_of(__ldg(f), __ldg(f + (long long)H * W), px, py, W, H);
        const Vec<V> g = vld<V>(gy + i * V);
        long long* base = acc + (((long long)b * H * W) * C4 + c) * V;
        const float w[4] = {k.wnw, k.wne, k.wsw, k.wse};
        const bool in[4] = {true, k.xin, k.yin, k.xin && k.yin};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!in[q]) continue;
            long long* p = base + ((long long)(k.y0 + (q >> 1)) * W + k.x0 + (q & 1)) * C4 * V;
#pragma unroll
            for (int j = 0; j < V; ++j) fx_add(p + j, w[q] * g.v[j], scale);
        }
    }
}
__global__ void fixed_to_float_kernel(const longlong2* __restrict__ acc, float2* __restrict__ out, const unsigned* __restrict__ absmax,
                                      long long n2) {
    const float inv = 1.f / fixed_scale(*absmax);
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n2; i += (long long)gridDim.x * NT) {
        const longlong2 a = acc[i];
        out[i] = make_float2((float)a.x * inv, (float)a.y * inv);
    }
}

// ---- PReLU -------------------------------------------------------------------------------------------------------------------
__global__ void prelu_cl_kernel(const float4* __restrict__ x, const float4* __restrict__ res, const float4* __restrict__ slope,
                                float4* __restrict__ y, long long total, int C4) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        float4 v = __ldg(x + i);
        if (res) v = f4_add(v, __ldg(res + i));
        const float4 s = __ldg(slope + (int)(i % C4));
        y[i] = make_float4(v.x > 0.f ? v.x : s.x * v.x, v.y > 0.f ? v.y : s.y * v.y, v.z > 0.f ? v.z : s.z * v.z, v.w > 0.f ? v.w : s.w * v.w);
    }
}

// ---- pose matrix -------------------------------------------------------------------------------------------------------------
struct Rod {
    float x, y, z, sa, ca, C, a, inv;   // axis, sin / cos / 1 - cos of the angle, angle, 1 / (angle + 1e-7) as a divisor
};
__device__ __forceinline__ Rod rodrigues(const float* v, float (&R)[3][3]) {
    Rod q;
    float s = __fmul_rn(v[0], v[0]);
    s = __fadd_rn(s, __fmul_rn(v[1], v[1]));
    s = __fadd_rn(s, __fmul_rn(v[2], v[2]));
    q.a = sqrtf(s);
    const float den = __fadd_rn(q.a, 1e-7f);
    q.inv = den;
    q.x = __fdiv_rn(v[0], den);
    q.y = __fdiv_rn(v[1], den);
    q.z = __fdiv_rn(v[2], den);
    q.ca = cosf(q.a);
    q.sa = sinf(q.a);
    q.C = __fsub_rn(1.f, q.ca);
    const float xs = __fmul_rn(q.x, q.sa), ys = __fmul_rn(q.y, q.sa), zs = __fmul_rn(q.z, q.sa);
    const float xC = __fmul_rn(q.x, q.C), yC = __fmul_rn(q.y, q.C), zC = __fmul_rn(q.z, q.C);
    const float xyC = __fmul_rn(q.x, yC), yzC = __fmul_rn(q.y, zC), zxC = __fmul_rn(q.z, xC);
    R[0][0] = __fadd_rn(__fmul_rn(q.x, xC), q.ca);
    R[0][1] = __fsub_rn(xyC, zs);
    R[0][2] = __fadd_rn(zxC, ys);
    R[1][0] = __fadd_rn(xyC, zs);
    R[1][1] = __fadd_rn(__fmul_rn(q.y, yC), q.ca);
    R[1][2] = __fsub_rn(yzC, xs);
    R[2][0] = __fsub_rn(zxC, ys);
    R[2][1] = __fadd_rn(yzC, xs);
    R[2][2] = __fadd_rn(__fmul_rn(q.z, zC), q.ca);
    return q;
}

__global__ void pose_matrix_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, float* __restrict__ M, int B, int invert) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float v[3] = {aa[3 * b], aa[3 * b + 1], aa[3 * b + 2]}, t[3] = {tr[3 * b], tr[3 * b + 1], tr[3 * b + 2]}, R[3][3];
    rodrigues(v, R);
    float* m = M + 16 * b;
    for (int i = 0; i < 3; ++i) {
        if (!invert) {   // T . R: rotation block as is, last column = t
            for (int j = 0; j < 3; ++j) m[4 * i + j] = R[i][j];
            m[4 * i + 3] = t[i];
        } else {         // R^T . T(-t): k-ordered multiply / add, as torch's 4x4 matmul evaluates it
            for (int j = 0; j < 3; ++j) m[4 * i + j] = R[j][i];
            float acc = __fmul_rn(R[0][i], -t[0]);
            acc = __fadd_rn(acc, __fmul_rn(R[1][i], -t[1]));
            acc = __fadd_rn(acc, __fmul_rn(R[2][i], -t[2]));
            m[4 * i + 3] = acc;
        }
    }
    m[12] = m[13] = m[14] = 0.f;
    m[15] = 1.f;
}

__global__ void pose_matrix_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, const float* __restrict__ gM,
                                       float* __restrict__ gaa, float* __restrict__ gtr, int B, int invert) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float v[3] = {aa[3 * b], aa[3 * b + 1], aa[3 * b + 2]}, t[3] = {tr[3 * b], tr[3 * b + 1], tr[3 * b + 2]}, R[3][3], G[3][3], gt[3];
    const Rod q = rodrigues(v, R);
    const float* g = gM + 16 * b;
    if (!invert) {
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) G[i][j] = g[4 * i + j];
            gt[i] = g[4 * i + 3];
        }
    } else {
        // M[i][j] = R[j][i], M[i][3] = sum_k R[k][i] * (-t[k])
        for (int k = 0; k < 3; ++k) {
            float s = 0.f;
            for (int i = 0; i < 3; ++i) {
                G[k][i] = g[4 * i + k] + g[4 * i + 3] * (-t[k]);
                s += R[k][i] * g[4 * i + 3];
            }
            gt[k] = -s;
        }
    }
    const float x = q.x, y = q.y, z = q.z, sa = q.sa, ca = q.ca, C = q.C;
    const float s01 = G[0][1] + G[1][0], s02 = G[0][2] + G[2][0], s12 = G[1][2] + G[2][1];
    const float d01 = G[1][0] - G[0][1], d02 = G[0][2] - G[2][0], d12 = G[2][1] - G[1][2];
    const float g_ca = G[0][0] + G[1][1] + G[2][2];
    const float g_C = G[0][0] * x * x + G[1][1] * y * y + G[2][2] * z * z + s01 * x * y + s02 * z * x + s12 * y * z;
    const float g_sa = d01 * z + d02 * y + d12 * x;
    const float gx = 2.f * x * C * G[0][0] + s01 * y * C + s02 * z * C + d12 * sa;
    const float gy = 2.f * y * C * G[1][1] + s01 * x * C + s12 * z * C + d02 * sa;
    const float gz = 2.f * z * C * G[2][2] + s02 * x * C + s12 * y * C + d01 * sa;
    float g_a = -sa * g_ca + ca * g_sa + sa * g_C;
    const float den = q.inv;
    g_a -= (gx * v[0] + gy * v[1] + gz * v[2]) / (den * den);
    const float ga_over = q.a > 0.f ? g_a / q.a : 0.f;
    gaa[3 * b + 0] = gx / den + ga_over * v[0];
    gaa[3 * b + 1] = gy / den + ga_over * v[1];
    gaa[3 * b + 2] = gz / den + ga_over * v[2];
    gtr[3 * b + 0] = gt[0];
    gtr[3 * b + 1] = gt[1];
    gtr[3 * b + 2] = gt[2];
}

}  // namespace

cudaError_t resize_bilinear_fwd(const float* x, float* y, int B, int C, int Hi, int Wi, int Ho, int Wo, float sh, float sw, int align,
                                float mul_even, float mul_odd, int layout, cudaStream_t st) {
    if (layout == 1) {
        const int V = (C % 4 == 0) ? 4 : 2;
        const long long total = (long long)B * Ho * Wo * (C / V);
        return launch_pdl(V == 4 ? resize_fwd_cl_kernel<4> : resize_fwd_cl_kernel<2>, dim3(grid_for(total)), dim3(NT), 0, st, x, y, B, C / V,
                          Hi, Wi, Ho, Wo, sh, sw, align);
    }
    const long long total = (long long)B * C * Ho * Wo;
    return launch_pdl(resize_fwd_nchw_kernel, dim3(grid_for(total)), dim3(NT), 0, st, x, y, B, C, Hi, Wi, Ho, Wo, sh, sw, align, mul_even,
                      mul_odd);
}

cudaError_t resize_bilinear_bwd(const float* gy, float* gx, int B, int C, int Hi, int Wi, int Ho, int Wo, float sh, float sw, int align,
                                float mul_even, float mul_odd, int layout, cudaStream_t st) {
    if (layout == 1) {
        const int V = (C % 4 == 0) ? 4 : 2;
        const long long total = (long long)B * Hi * Wi * (C / V);
        return launch_pdl(V == 4 ? resize_bwd_cl_kernel<4> : resize_bwd_cl_kernel<2>, dim3(grid_for(total)), dim3(NT), 0, st, gy, gx, B, C / V,
                          Hi, Wi, Ho, Wo, sh, sw, align);
    }
    const long long total = (long long)B * C * Hi * Wi;
    return launch_pdl(resize_bwd_nchw_kernel, dim3(grid_for(total)), dim3(NT), 0, st, gy, gx, B, C, Hi, Wi, Ho, Wo, sh, sw, align, mul_even,
                      mul_odd);
}

cudaError_t flow_warp_fwd(const float* x, const float* flow, float* y, int B, int C, int H, int W, int layout, cudaStream_t st) {
    if (layout == 1) {
        const int V = (C % 4 == 0) ? 4 : 2;
        const long long total = (long long)B * H * W * (C / V);
        return launch_pdl(V == 4 ? flow_warp_fwd_cl_kernel<4> : flow_warp_fwd_cl_kernel<2>, dim3(grid_for(total)), dim3(NT), 0, st, x, flow, y, B,
                          C / V, H, W);
    }
    return launch_pdl(flow_warp_fwd_nchw_kernel, dim3(grid_for((long long)B * H * W)), dim3(NT), 0, st, x, flow, y, B, C, H, W);
}

size_t flow_warp_bwd_workspace_bytes(int B, int C, int H, int W) { return 16 + (size_t)B * C * H * W * sizeof(long long); }

cudaError_t flow_warp_bwd(const float* gy, const float* flow, float* gx, int B, int C, int H, int W, void* workspace, size_t workspace_bytes,
                          cudaStream_t st) {
    const size_t need = flow_warp_bwd_workspace_bytes(B, C, H, W);
    if (workspace_bytes < need) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
    if (e != cudaSuccess) return e;
    unsigned* amax = reinterpret_cast<unsigned*>(workspace);
    long long* acc = reinterpret_cast<long long*>(reinterpret_cast<char*>(workspace) + 16);
    const int V = (C % 4 == 0) ? 4 : 2;
    const long long n2 = (long long)B * H * W * (C / 2), nv = (long long)B * H * W * (C / V);
    absmax_kernel<<<grid_for(n2), NT, 0, st>>>((const float2*)gy, n2, amax);
    if (V == 4) flow_warp_scatter_kernel<4><<<grid_for(nv), NT, 0, st>>>(gy, flow, acc, amax, B, C / 4, H, W);
    else flow_warp_scatter_kernel<2><<<grid_for(nv), NT, 0, st>>>(gy, flow, acc, amax, B, C / 2, H, W);
    fixed_to_float_kernel<<<grid_for(n2), NT, 0, st>>>((const longlong2*)acc, (float2*)gx, amax, n2);
    return cudaGetLastError();
}

cudaError_t prelu_cl_fwd(const float* x, const float* res, const float* slope, float* y, long long P, int C, cudaStream_t st) {
    const long long total = P * (C / 4);
    return launch_pdl(prelu_cl_kernel, dim3(grid_for(total)), dim3(NT), 0, st, (const float4*)x, (const float4*)res, (const float4*)slope,
                      (float4*)y, total, C / 4);
}

cudaError_t pose_matrix_fwd(const float* aa, const float* tr, float* M, int B, int invert, cudaStream_t st) {
    pose_matrix_fwd_kernel<<<(B + 63) / 64, 64, 0, st>>>(aa, tr, M, B, invert);
    return cudaGetLastError();
}
cudaError_t pose_matrix_bwd(const float* aa, const float* tr, const float* gM, float* gaa, float* gtr, int B, int invert, cudaStream_t st) {
    pose_matrix_bwd_kernel<<<(B + 63) / 64, 64, 0, st>>>(aa, tr, gM, gaa, gtr, B, invert);
    return cudaGetLastError();
}

}  // namespace mvf
