// F1 backward: gradient of the fused photometric loss (f1_fwd.cu) w.r.t. disp and P0/P1, recomputing the
// view synthesis instead of storing it (44 B/px algorithmic: re-read 40, write grad_disp 4; the argmin map
// written by the forward adds 1 B/px).  Formulas: SURVEY.md 9.3; checked against oracle/f1_oracle.c and the
// reference's autograd (tests/golden).
//
// Per 32x16 tile:
//   phase 1  warped sources + target at tile + 2px halo (reflection-padded) -> smem; bilinear
//            derivatives of the interior pixels; selection weights gs_k(p) = gout*share_k*mask/n at tile + 1
//   phase 2  SSIM window statistics at tile + 1 (separable, packed FFMA2) -> per-window adjoint
//            coefficients (C_mu, C_xx, C_xy) per channel and candidate
//   phase 3  3x3 adjoint stencil of the coefficients (weights 0/1/2 encode the reflection multiplicity)
//            -> d loss / d warped -> grid -> camera point -> depth -> disp, and the [3,4] gradient of P
#include "f1.cuh"

namespace mvf {

namespace {

constexpr int TW = 32, TH = 16;
constexpr int W2 = TW + 4, H2 = TH + 4;  // tile + 2
constexpr int W1 = TW + 2, H1 = TH + 2;  // tile + 1
constexpr int NT = 256;
constexpr int R2 = 6;                    // window rows per thread in phase 2 (H1 = 3 * R2)
constexpr int NCP = W1 / 2;              // 17 column pairs in phase 2
constexpr int NITEM2 = 3 * NCP * (H1 / R2);

// Geometry of the recompute.  Exact (IEEE divisions, bit-identical sampling weights to the forward) by
// default; -DMVF_BWD_FAST_GEOMETRY=1 switches to MUFU reciprocals (coordinates differ by ~1e-4 px).
#ifndef MVF_BWD_FAST_GEOMETRY
#define MVF_BWD_FAST_GEOMETRY 0
#endif
__device__ __forceinline__ void project_bwd_tap(float depth, const float c[3], const float* __restrict__ P, const Geo& g,
                                                Tap& t, float X[3], float pr[3], float& rz) {
#if MVF_BWD_FAST_GEOMETRY
    project_tap_fast(depth, c, P, g, t, X, pr, rz);
#else
    project_tap(depth, c, P, g, t, X, pr);
    rz = __fdividef(1.0f, pr[2] + 1e-7f);
#endif
}
__device__ __forceinline__ float depth_bwd(float d, float min_disp, float range) {
#if MVF_BWD_FAST_GEOMETRY
    return __fdividef(1.0f, min_disp + range * d);
#else
    return disp_to_depth(d, min_disp, range);
#endif
}

struct __align__(16) BwdSmem {
    float2 T2[3][H2][W2];     // target {t,t}
    float2 Wp[3][H2][W2];     // warped {warp0, warp1}
    float2 CF[3][3][H1][W1];  // [channel][field mu,xx,xy][row][col] adjoint coefficients {k=0, k=1}
    float4 DXY[3][TH][TW];    // {dwarp0/dix, dwarp1/dix, dwarp0/diy, dwarp1/diy}
    float2 GS[H1][W1];        // selection weight of each window {k=0, k=1}
    float D[H1][W1];          // disparity at tile + 1
    float cst[36];
    float uni[8];             // CTA-uniform scalars computed once by one thread: {gn, cxn, cyn, den, shift}
};
static_assert(sizeof(float) * 24 * NT <= sizeof(float2) * 3 * 3 * H1 * W1, "reduction scratch aliases CF");

struct HB {
    float2 t, tt, w, ww, wt;
};

__device__ __forceinline__ void hsum_row_b(const BwdSmem& sm, int c, int row, int cp, HB h[2]) {
    const float4* tp = reinterpret_cast<const float4*>(&sm.T2[c][row][2 * cp]);
    const float4* wp = reinterpret_cast<const float4*>(&sm.Wp[c][row][2 * cp]);
    float4 ta = tp[0], tb = tp[1], wa = wp[0], wb = wp[1];
    float2 t0 = make_float2(ta.x, ta.y), t1 = make_float2(ta.z, ta.w), t2 = make_float2(tb.x, tb.y),
           t3 = make_float2(tb.z, tb.w);
    float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w), w2 = make_float2(wb.x, wb.y),
           w3 = make_float2(wb.z, wb.w);
    float2 m = add2(t1, t2);
    h[0].t = add2(m, t0);
    h[1].t = add2(m, t3);
    m = fma2(t1, t1, mul2(t2, t2));
    h[0].tt = fma2(t0, t0, m);
    h[1].tt = fma2(t3, t3, m);
    m = add2(w1, w2);
    h[0].w = add2(m, w0);
    h[1].w = add2(m, w3);
    m = fma2(w1, w1, mul2(w2, w2));
    h[0].ww = fma2(w0, w0, m);
    h[1].ww = fma2(w3, w3, m);
    m = fma2(w1, t1, mul2(w2, t2));
    h[0].wt = fma2(w0, t0, m);
    h[1].wt = fma2(w3, t3, m);
}

// adjoint coefficients of one SSIM window for the packed pair of warped candidates
__device__ __forceinline__ void emit_coeff(BwdSmem& sm, int c, int row, int col, const HB& a, const HB& b, const HB& cu,
                                           float cS) {
    const float2 k9 = f2(1.0f / 9.0f);
    const float C1 = 0.0001f, C2 = 0.0009f;
    float2 vt = add2(add2(a.t, b.t), cu.t), vtt = add2(add2(a.tt, b.tt), cu.tt);
    float2 vw = add2(add2(a.w, b.w), cu.w), vww = add2(add2(a.ww, b.ww), cu.ww), vwt = add2(add2(a.wt, b.wt), cu.wt);
    float2 my = mul2(vt, k9), mx = mul2(vw, k9);
    float2 my2 = mul2(my, my), mx2 = mul2(mx, mx), mxmy = mul2(mx, my);
    float2 vy = fma2(vtt, k9, -my2), vx = fma2(vww, k9, -mx2), vxy = fma2(vwt, k9, -mxmy);
    float2 A = fma2(f2(2.0f), mxmy, f2(C1)), Bn = fma2(f2(2.0f), vxy, f2(C2));
    float2 Cd = add2(add2(mx2, my2), f2(C1)), Dd = add2(add2(vx, vy), f2(C2));
    float2 den = mul2(Cd, Dd);
    float2 inv = make_float2(__fdividef(1.0f, den.x), __fdividef(1.0f, den.y));
    float2 S = mul2(mul2(A, Bn), inv);
    // clamp((1-S)/2, 0, 1) passes gradient on 0 <= raw <= 1  <=>  -1 <= S <= 1
    float2 gs = sm.GS[row][col];
    float2 coef = mul2(gs, f2(cS * (-0.5f) / 9.0f));
    coef.x = (S.x >= -1.0f && S.x <= 1.0f) ? coef.x : 0.0f;
    coef.y = (S.y >= -1.0f && S.y <= 1.0f) ? coef.y : 0.0f;
    // g_mu = 2my(Bn-A)inv - 2mx S (D-Cd) inv ; g_xx = -S Cd inv ; g_xy = 2 A inv
    float2 g_mu = mul2(sub2(mul2(mul2(my, f2(2.0f)), sub2(Bn, A)), mul2(mul2(mul2(mx, f2(2.0f)), S), sub2(Dd, Cd))), inv);
    float2 g_xx2 = mul2(mul2(mul2(S, Cd), inv), f2(-2.0f));
    float2 g_xy = mul2(mul2(A, f2(2.0f)), inv);
    sm.CF[c][0][row][col] = mul2(coef, g_mu);
    sm.CF[c][1][row][col] = mul2(coef, g_xx2);
    sm.CF[c][2][row][col] = mul2(coef, g_xy);
}

// reflection multiplicity of window p = q + d as seen from pixel q (0 outside the image, 2 at the fold)
__device__ __forceinline__ float mult_lo(int q) { return q == 0 ? 0.0f : (q == 1 ? 2.0f : 1.0f); }
__device__ __forceinline__ float mult_hi(int q, int n) { return q == n - 1 ? 0.0f : (q == n - 2 ? 2.0f : 1.0f); }

__device__ __forceinline__ float sgnf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

__global__ void __launch_bounds__(NT, 2) f1_bwd_kernel(const F1Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int b = blockIdx.z, tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int H = a.H, W = a.W, B = a.B;
    const size_t HW = (size_t)H * W;
    const bool avg = (a.flags & F_AVG_REPROJECTION) != 0, am = !(a.flags & F_DISABLE_AUTOMASKING);
    const bool nossim = (a.flags & F_NO_SSIM) != 0;
    const float cS = nossim ? 0.0f : 0.85f / 3.0f, cL = nossim ? 1.0f / 3.0f : 0.15f / 3.0f;
    const float gout = a.gout ? __ldg(a.gout) : 1.0f;
    const Geo g = make_geo(H, W);

    const float* dispb = a.disp + (size_t)b * HW;
    const float* tgtb = a.tgt + (size_t)b * 3 * HW;
    const int HWi = H * W;
    // disparity and target of a halo position (reflection-padded, clamped): the loads of the NEXT position of this thread are issued
    // before the projections and gathers of the current one (and the first ones before the barrier below), so that their latency is
    // not the first thing every iteration of phase 1 waits for
    auto pix_of = [&](int p) {
        const int hy = p / W2, hx = p - hy * W2;
        const int y = clampi(reflect1(ty0 - 2 + hy, H), 0, H - 1), x = clampi(reflect1(tx0 - 2 + hx, W), 0, W - 1);
        return y * W + x;
    };
    float d_nx, tv_nx[3];
    {
        const int i = pix_of(tid);   // NT <= H2 * W2: every thread has a first position
        d_nx = __ldg(dispb + i);
#pragma unroll
        for (int c = 0; c < 3; ++c) tv_nx[c] = __ldg(tgtb + (i + c * HWi));
    }
    if (tid < 12) sm.cst[tid] = a.inv_K[16 * b + tid];
    else if (tid < 24) sm.cst[tid] = a.P0[12 * b + tid - 12];
    else if (tid < 36) sm.cst[tid] = a.P1[12 * b + tid - 24];
    else if (tid == 36) {
        // normalisation constants (double conversions, divisions and the three dependent loads of the forward's statistics) once per
        // CTA instead of once per thread: they were 11 % of the kernel's stall samples (profiles/r1_f1_bwd_ncu_lines.txt, line 273)
        const float mean = a.stats[4 * b + 0], Sx = a.stats[4 * b + 1], Sy = a.stats[4 * b + 2];
        const float den = mean + 1e-7f;
        const float cxn = gout * a.smooth_w / (float)((double)B * H * (W - 1));
        const float cyn = gout * a.smooth_w / (float)((double)B * (H - 1) * W);
        const float dotb = cxn * Sx + cyn * Sy;  // sum_q g_nd[q] * disp[q]
        sm.uni[0] = gout / (float)((double)B * (double)HW);
        sm.uni[1] = cxn;
        sm.uni[2] = cyn;
        sm.uni[3] = den;
        sm.uni[4] = dotb / (den * den * (float)HW);
    }
    __syncthreads();

    const float* s0b = a.src0 + (size_t)b * 3 * HW;
    const float* s1b = a.src1 + (size_t)b * 3 * HW;
    const uint8_t* __restrict__ idxb = a.idx + (size_t)b * HW;
    const float* __restrict__ mkb = a.mask ? a.mask + (size_t)b * HW : nullptr;

    // ---------------- phase 1 ---------------------------------------------------------------------------------
    {
        const int first_rep = am ? (avg ? 1 : 2) : 0;
        const float gn = sm.uni[0];
        for (int p = tid; p < H2 * W2; p += NT) {
            int hy = p / W2, hx = p - hy * W2;
            int ry = ty0 - 2 + hy, rx = tx0 - 2 + hx;
            int y = clampi(reflect1(ry, H), 0, H - 1), x = clampi(reflect1(rx, W), 0, W - 1);
            const int i = y * W + x;
            const float d = d_nx;
            float tv[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) tv[c] = tv_nx[c];
            if (p + NT < H2 * W2) {
                const int in = pix_of(p + NT);
                d_nx = __ldg(dispb + in);
#pragma unroll
                for (int c = 0; c < 3; ++c) tv_nx[c] = __ldg(tgtb + (in + c * HWi));
            }
            float depth = depth_bwd(d, a.min_disp, a.disp_range);
            float cr[3], X[3], pr[3], rz;
            cam_ray(sm.cst, (float)x, (float)y, cr);
            Tap t0, t1;
            float w0[3], w1[3], dx0[3], dy0[3], dx1[3], dy1[3];
            project_bwd_tap(depth, cr, sm.cst + 12, g, t0, X, pr, rz);
            gather3_grad(s0b, HWi, W, corner_of(t0, H, W), w0, dx0, dy0);
            project_bwd_tap(depth, cr, sm.cst + 24, g, t1, X, pr, rz);
            gather3_grad(s1b, HWi, W, corner_of(t1, H, W), w1, dx1, dy1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                sm.T2[c][hy][hx] = make_float2(tv[c], tv[c]);
                sm.Wp[c][hy][hx] = make_float2(w0[c], w1[c]);
            }
            const bool in1 = hy >= 1 && hy <= H1 && hx >= 1 && hx <= W1;  // tile + 1
            if (in1) {
                const bool inimg = (ry == y) && (rx == x) && ry >= 0 && rx >= 0;  // raw position is a real pixel
                float2 gs = make_float2(0.f, 0.f);
                if (inimg) {
                    int sel = idxb[i];
                    float m = mkb ? __ldg(mkb + i) : 1.0f;
                    float sh0, sh1;
                    if (avg) sh0 = sh1 = (!am || sel == first_rep) ? 0.5f : 0.0f;
                    else { sh0 = (sel == first_rep) ? 1.0f : 0.0f; sh1 = (sel == first_rep + 1) ? 1.0f : 0.0f; }
                    gs = make_float2(gn * m * sh0, gn * m * sh1);
                }
                sm.GS[hy - 1][hx - 1] = gs;
                sm.D[hy - 1][hx - 1] = d;
            }
            if (hy >= 2 && hy < 2 + TH && hx >= 2 && hx < 2 + TW) {
#pragma unroll
                for (int c = 0; c < 3; ++c) sm.DXY[c][hy - 2][hx - 2] = make_float4(dx0[c], dx1[c], dy0[c], dy1[c]);
            }
        }
    }
    __syncthreads();

    // ---------------- phase 2: window statistics at tile + 1 -> adjoint coefficients ---------------------------
    if (tid < NITEM2) {
        const int c = tid / (NCP * (H1 / R2)), rem = tid - c * (NCP * (H1 / R2));
        const int rg = rem / NCP, cp = rem - rg * NCP;
        HB r0[2], r1[2], cu[2];
        hsum_row_b(sm, c, rg * R2 + 0, cp, r0);
        hsum_row_b(sm, c, rg * R2 + 1, cp, r1);
#pragma unroll
        for (int s = 2; s < R2 + 2; ++s) {
            hsum_row_b(sm, c, rg * R2 + s, cp, cu);
            emit_coeff(sm, c, rg * R2 + s - 2, 2 * cp, r0[0], r1[0], cu[0], cS);
            emit_coeff(sm, c, rg * R2 + s - 2, 2 * cp + 1, r0[1], r1[1], cu[1], cS);
#pragma unroll
            for (int j = 0; j < 2; ++j) { r0[j] = r1[j]; r1[j] = cu[j]; }
        }
    }
    __syncthreads();

    // ---------------- phase 3: adjoint stencil, chain rule to disp and P ---------------------------------------
    float* __restrict__ gdb = a.g_disp + (size_t)b * HW;
    float gP[24];
#pragma unroll
    for (int q = 0; q < 24; ++q) gP[q] = 0.f;
    {
        const int row = tid >> 4, cp = tid & 15;  // 16 rows x 16 column pairs = 256 threads
        const int y = ty0 + row;
        const float wu = mult_lo(y), wd = mult_hi(y, H);
        float2 gix[2], giy[2];
        gix[0] = gix[1] = giy[0] = giy[1] = make_float2(0.f, 0.f);
        float wl[2], wr[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int x = tx0 + 2 * cp + j;
            wl[j] = mult_lo(x);
            wr[j] = mult_hi(x, W);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float2 G[3][2];  // [field][output column]
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                G[f][0] = G[f][1] = make_float2(0.f, 0.f);
#pragma unroll
                for (int dr = 0; dr < 3; ++dr) {
                    const float4* cpz = reinterpret_cast<const float4*>(&sm.CF[c][f][row + dr][2 * cp]);
                    float4 ca = cpz[0], cb = cpz[1];
                    float2 c0 = make_float2(ca.x, ca.y), c1 = make_float2(ca.z, ca.w), c2 = make_float2(cb.x, cb.y),
                           c3 = make_float2(cb.z, cb.w);
                    float wy = dr == 0 ? wu : (dr == 2 ? wd : 1.0f);
                    float2 h0 = fma2(f2(wr[0]), c2, fma2(f2(wl[0]), c0, c1));
                    float2 h1 = fma2(f2(wr[1]), c3, fma2(f2(wl[1]), c1, c2));
                    G[f][0] = fma2(f2(wy), h0, G[f][0]);
                    G[f][1] = fma2(f2(wy), h1, G[f][1]);
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float2 t = sm.T2[c][row + 2][2 * cp + j + 2];
                float2 w = sm.Wp[c][row + 2][2 * cp + j + 2];
                float2 gsq = sm.GS[row + 1][2 * cp + j + 1];
                float2 ga = fma2(t, G[2][j], fma2(w, G[1][j], G[0][j]));
                float2 df = sub2(t, w);
                ga.x = fmaf(gsq.x * cL, -sgnf(df.x), ga.x);
                ga.y = fmaf(gsq.y * cL, -sgnf(df.y), ga.y);
                float4 dxy = sm.DXY[c][row][2 * cp + j];
                gix[j] = fma2(ga, make_float2(dxy.x, dxy.y), gix[j]);
                giy[j] = fma2(ga, make_float2(dxy.z, dxy.w), giy[j]);
            }
        }
        // smoothness constants (train.py:1044-1049, layers.py:231-242)
        const float cxn = sm.uni[1], cyn = sm.uni[2], den = sm.uni[3], shift = sm.uni[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = 2 * cp + j, x = tx0 + col;
            if (y >= H || x >= W) continue;
            const float d = sm.D[row + 1][col + 1];
            const float depth = depth_bwd(d, a.min_disp, a.disp_range);
            float cr[3];
            cam_ray(sm.cst, (float)x, (float)y, cr);
            float gdepth = 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float* P = sm.cst + 12 + 12 * k;
                Tap t;
                float X[3], pr[3], rz;
                project_bwd_tap(depth, cr, P, g, t, X, pr, rz);
                float gx = k == 0 ? gix[j].x : gix[j].y, gy = k == 0 ? giy[j].x : giy[j].y;
                if (!(t.ixr > 0.0f && t.ixr < g.wm1)) gx = 0.f;
                if (!(t.iyr > 0.0f && t.iyr < g.hm1)) gy = 0.f;
                float gp[3];
                gp[0] = gx * rz;
                gp[1] = gy * rz;
                gp[2] = -(gx * pr[0] + gy * pr[1]) * rz * rz;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    gP[12 * k + 4 * r + 0] = fmaf(gp[r], X[0], gP[12 * k + 4 * r + 0]);
                    gP[12 * k + 4 * r + 1] = fmaf(gp[r], X[1], gP[12 * k + 4 * r + 1]);
                    gP[12 * k + 4 * r + 2] = fmaf(gp[r], X[2], gP[12 * k + 4 * r + 2]);
                    gP[12 * k + 4 * r + 3] += gp[r];
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    float gX = gp[0] * P[q] + gp[1] * P[4 + q] + gp[2] * P[8 + q];
                    gdepth = fmaf(gX, cr[q], gdepth);
                }
            }
            float gd = -gdepth * depth * depth * a.disp_range;
            // edge-aware smoothness
            float gnd = 0.f;
            const float tc0 = sm.T2[0][row + 2][col + 2].x, tc1 = sm.T2[1][row + 2][col + 2].x,
                        tc2 = sm.T2[2][row + 2][col + 2].x;
            if (x + 1 < W) {
                float gi = fabsf(tc0 - sm.T2[0][row + 2][col + 3].x) + fabsf(tc1 - sm.T2[1][row + 2][col + 3].x) +
                           fabsf(tc2 - sm.T2[2][row + 2][col + 3].x);
                gnd += cxn * sgnf(d - sm.D[row + 1][col + 2]) * __expf(-(gi / 3.0f));
            }
            if (x >= 1) {
                float gi = fabsf(tc0 - sm.T2[0][row + 2][col + 1].x) + fabsf(tc1 - sm.T2[1][row + 2][col + 1].x) +
                           fabsf(tc2 - sm.T2[2][row + 2][col + 1].x);
                gnd -= cxn * sgnf(sm.D[row + 1][col] - d) * __expf(-(gi / 3.0f));
            }
            if (y + 1 < H) {
                float gi = fabsf(tc0 - sm.T2[0][row + 3][col + 2].x) + fabsf(tc1 - sm.T2[1][row + 3][col + 2].x) +
                           fabsf(tc2 - sm.T2[2][row + 3][col + 2].x);
                gnd += cyn * sgnf(d - sm.D[row + 2][col + 1]) * __expf(-(gi / 3.0f));
            }
            if (y >= 1) {
                float gi = fabsf(tc0 - sm.T2[0][row + 1][col + 2].x) + fabsf(tc1 - sm.T2[1][row + 1][col + 2].x) +
                           fabsf(tc2 - sm.T2[2][row + 1][col + 2].x);
                gnd -= cyn * sgnf(sm.D[row][col + 1] - d) * __expf(-(gi / 3.0f));
            }
            gd += gnd / den - shift;
            gdb[y * W + x] = gd;
        }
    }
    // ---------------- block reduction of grad_P (scratch aliases CF), fixed-point atomics ---------------------
    __syncthreads();
    float* red = reinterpret_cast<float*>(&sm.CF[0][0][0][0]);
#pragma unroll
    for (int q = 0; q < 24; ++q) red[q * NT + tid] = gP[q];
    __syncthreads();
    long long* acc = ws_bwd_acc(a.ws, B);
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int q = warp; q < 24; q += NT / 32) {
            float v = 0.f;
#pragma unroll
            for (int s = 0; s < NT / 32; ++s) v += red[q * NT + s * 32 + lane];
            double dv = warp_sum((double)v);
            if (lane == 0) {
                int k = q / 12, e = q - 12 * k;
                atomicAdd(reinterpret_cast<unsigned long long*>(acc + ((size_t)k * B + b) * 12 + e),
                          (unsigned long long)to_fix(dv));
            }
        }
    }
    // (no fence, no last-CTA ticket: the fixed-point sums are turned into the outputs by f1_bwd_finalize_kernel, the next launch on the
    // stream.  The fence -- it waits for this warp's atomics to reach L2 -- and the ticket cost every one of the 2880 CTAs two block
    // barriers and ~6 % of the kernel's stall samples, profiles/r2_f1_bwd_final.keys.txt.)
}

// [2][B][12] fixed-point sums -> grad_P0 / grad_P1, and the accumulators are zeroed for the next launch (self-cleaning workspace)
__global__ void f1_bwd_finalize_kernel(const F1Args a) {
    long long* acc = ws_bwd_acc(a.ws, a.B);
    const int B = a.B;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 24 * B; i += gridDim.x * blockDim.x) {
        const double val = from_fix(acc[i]);
        acc[i] = 0;
        const int k = i / (12 * B), e = i - k * 12 * B;
        (k == 0 ? a.g_P0 : a.g_P1)[e] = (float)val;
    }
}

}  // namespace

cudaError_t launch_f1_backward(const F1Args& a, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(f1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(BwdSmem));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, a.B);
    f1_bwd_kernel<<<grid, NT, sizeof(BwdSmem), stream>>>(a);
    f1_bwd_finalize_kernel<<<(24 * a.B + 255) / 256, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace mvf
