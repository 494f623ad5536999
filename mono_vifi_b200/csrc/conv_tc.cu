// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, TF32 inputs, fp32 accumulation in TMEM) on
// channels-last fp32 activations (NCHW-shaped tensors in torch.channels_last memory format) -- no im2col buffer.
//
//   fprop :  y[b,oy,ox,co] = sum_{kh,kw,ci} x[b, oy*s+kh-pad, ox*s+kw-pad, ci] * w[co,ci,kh,kw]     (stride s = 1 or 2)
//   dgrad :  (s = 1) the same kernel on grad_out with the flipped / transposed filter bank and pad' = k-1-pad
//
// GEMM view per CTA: D[128 pixels, N_TILE couts] += A[128 pixels, 32 channels] . B[32 channels, N_TILE] for each
// (filter tap, 32-channel block).  The 128 pixels are a TW x TH patch of one image (TW * TH = 128, chosen per layer),
// so the A operand of a tap is one TMA box of the input, shifted by (kh-pad, kw-pad) and traversed with the
// convolution stride; pixels outside the image (zero padding) and channels past Cin are zero-filled by the TMA unit.
// Both operands are K-major with the 128-byte swizzle (A: one row per pixel, B: one row per cout of the pre-packed
// filter bank).  [Why channels-last: with NCHW the shifted box would start at a 4-byte offset in the contiguous
// dimension, which TMA rejects (illegal instruction, measured), and MN-major TF32 operands need the 32-byte-unit
// swizzle; profiles/r1_conv_bringup.md.]
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected thread), warps 2..5 =
// epilogue (TMEM -> registers -> bias / activation -> 128-byte-per-thread NHWC stores).
//
// Replaces the cuDNN calls behind nn.Conv2d in the reference's networks (layers.py:131,146; monodepth2.py; posenet.py).
#include "conv_tc.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "pdl.cuh"
#include "tc_common.cuh"

namespace mvf {
namespace tc {

namespace {

constexpr int TILE_M = 128;                         // output pixels per CTA (= UMMA M)
constexpr int BLOCK_K = 32, KGROUPS = BLOCK_K / 8;  // channels per pipeline stage, UMMA K = 8 (tf32)
constexpr int A_STAGE_BYTES = TILE_M * BLOCK_K * 4; // 16 KB
constexpr int NTHREADS = 192;
constexpr int IGEMM_THREADS = 224;  // conv_igemm_kernel: warp 6 is the second MMA issuer (odd pipeline iterations)
constexpr int PATCH_THREADS = 224;  // conv_patch_kernel: warp 6 is the second MMA issuer (stacked tile mt = 1)

template <int N_TILE>
struct Cfg {
    static constexpr int B_STAGE_BYTES = N_TILE * BLOCK_K * 4;
    static constexpr int STAGES = (N_TILE >= 256) ? 4 : ((N_TILE >= 128) ? 3 : 4);
    static constexpr int TMEM_COLS = (2 * N_TILE) < 32 ? 32 : (2 * N_TILE);  // conv_igemm_kernel: one accumulator per MMA issuer
    static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct ConvArgs {
    float* y;
    const float* bias;
    long long y_sB, y_sH, y_sW;  // output strides in elements (channel stride 1)
    int Cout, Ho, Wo;
    int KH, KW, pad, stride, stride_x;  // row / column stride of the convolution
    int tw_log2;                 // tile = (1 << tw_log2) x (128 >> tw_log2) output pixels
    int tiles_x, tiles_y;
    int n_cblk;                  // ceil(Cin / 32)
    int act;                     // 0 none, 1 relu, 2 elu, 3 prelu (per-channel `slope`)
    const float* slope;
    float* dbg;                  // debug: raw copy of pipeline stage 0 (A then B) of CTA (0,0); null in production
    int single_issuer;           // 1: one MMA issuer (MVF_IGEMM_ISSUERS=1, A/B timing and bisection)
};

// ---- epilogue arithmetic: v[0..CH) += bias[n_first ..], then the activation ---------------------------------------
// ELU's exp(t) - 1 for t <= 0: ex2.approx based, with the cubic Taylor polynomial where cancellation would cost accuracy
// (|t| < 1/16: polynomial error < 7e-7 relative; elsewhere the subtraction loses at most ~4 bits of a 2-ulp exp) --
// far inside the TF32 product error of the convolution it follows, and 5 instructions instead of expm1f's ~35.
__device__ __forceinline__ float elu_neg(float t) {
    const float e = __expf(t) - 1.0f;
    const float q = t * fmaf(t, fmaf(t, 0.16666667f, 0.5f), 1.0f) + t * t * t * t * 0.041666668f;
    return t > -0.0625f ? q : e;
}

// Branches are on kernel-uniform values only and sit OUTSIDE the element loops (the per-element form cost ~30
// instructions per value: it was 40 % of the stall samples of the 16-channel decoder layers).
template <int CH>
__device__ __forceinline__ void bias_act(float (&v)[CH], const float* __restrict__ bias, int n_first, int Cout, int act,
                                         const float* __restrict__ slope = nullptr) {
    if (bias) {
        if (n_first + CH <= Cout && ((reinterpret_cast<uintptr_t>(bias + n_first) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n_first + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
                if (n_first + j < Cout) v[j] += __ldg(bias + n_first + j);
        }
    }
    if (act == 1) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = v[j] > 0.f ? v[j] : elu_neg(v[j]);
    } else if (act == 3) {   // nn.PReLU(Cout) of the VFI network (IFRNet.py:121-125)
#pragma unroll
        for (int j = 0; j < CH; ++j)
            if (n_first + j < Cout) v[j] = v[j] > 0.f ? v[j] : __ldg(slope + n_first + j) * v[j];
    }
}

template <int N_TILE>
__global__ void __launch_bounds__(IGEMM_THREADS) conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                              const __grid_constant__ CUtensorMap mapB, const ConvArgs p) {
    using C = Cfg<N_TILE>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment: the 128B-swizzle atoms of TMA and UMMA are defined on absolute shared-memory address bits
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* smA = smem;
    unsigned char* smB = smem + C::STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smB + C::STAGES * C::B_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full_bar = empty_bar + C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x;
    const int tx = m % p.tiles_x, ty = (m / p.tiles_x) % p.tiles_y, b = m / (p.tiles_x * p.tiles_y);
    const int TW = 1 << p.tw_log2, TH = TILE_M >> p.tw_log2;
    const int x0 = tx * TW, y0 = ty * TH, n0 = blockIdx.y * N_TILE;
    const int n_iters = p.KH * p.KW * p.n_cblk;
    // two MMA issuers (see conv_patch_kernel): warp 1 multiplies the even pipeline iterations into accumulator columns [0, N), warp 6
    // the odd ones into [N, 2N); the epilogue adds the two halves (a fixed order, so results stay bitwise repeatable)
    const int n_issuers = (n_iters >= 2 && !p.single_issuer) ? 2 : 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], n_issuers);   // the owner's commit + the other issuer's "I have seen this phase"
        }
        mbar_init(tmem_full_bar, n_issuers);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    pdl_sync();  // prologue done: wait for the producer of our inputs, let the next kernel start its own prologue

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % C::STAGES;
                const uint32_t ph = (it / C::STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int tap = it / p.n_cblk, cb = it - tap * p.n_cblk;
                const int kh = tap / p.KW, kw = tap - kh * p.KW;
                mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + C::B_STAGE_BYTES);
                // A: dims (c, x, y, b), box (32, TW, TH, 1) traversed with the conv stride -> smem [pixel][32 c]
                tma_load_4d(smA + s * A_STAGE_BYTES, &mapA, &full_bar[s], cb * BLOCK_K, x0 * p.stride_x + kw - p.pad,
                            y0 * p.stride + kh - p.pad, b);
                // B: dims (k within block, cout, tap * n_cblk + cb), box (32, N_TILE, 1) -> smem [cout][32 k]
                tma_load_3d(smB + s * C::B_STAGE_BYTES, &mapB, &full_bar[s], 0, n0, it);
            }
        }
    } else if (warp == 1 || warp == 6) {
        // ===== MMA issuers =====
        const int issuer = warp == 1 ? 0 : 1;
        if (issuer < n_issuers && elect_one()) {
            constexpr uint32_t idesc = make_idesc_tf32(TILE_M, N_TILE, /*A K-major*/ 0, /*B K-major*/ 0);
            // K-major SW128: rows of 32 k (128 B), 8-row groups 1024 B apart (SBO); k-group kg starts 32 B (2 units) in
            const uint64_t desc_hi = make_smem_desc(0, 16, 1024, SWZ_128B);
            const uint64_t d_up = desc_hi & 0xFFFFFFFF00000000ull;
            const uint32_t lo_const = (uint32_t)desc_hi;
            const uint32_t a0 = (smem_u32(smA) >> 4) | lo_const, b0 = (smem_u32(smB) >> 4) | lo_const;
            const uint32_t d = tmem_d + (uint32_t)(issuer * N_TILE);
            int s = 0;
            uint32_t ph = 0, accum = 0;
            uint32_t a_lo = a0, b_lo = b0;
            // Both issuers wait on EVERY stage barrier, in order, and both release the stage (the owner of the iteration with the
            // commit of its MMAs, the other with a plain arrive): try_wait.parity only tells the current phase from the preceding
            // one, so an issuer must never fall a whole ring generation behind the barriers it polls.
            for (int it = 0; it < n_iters; ++it) {
                mbar_wait(&full_bar[s], ph);
                if ((it & 1) != issuer && n_issuers == 2) {
                    mbar_arrive(&empty_bar[s]);
                } else {
                    tc_fence_after();
                    umma_tf32(d, d_up | (uint64_t)a_lo, d_up | (uint64_t)b_lo, idesc, accum);
#pragma unroll
                    for (int kg = 1; kg < KGROUPS; ++kg)
                        umma_tf32_acc(d, d_up | (uint64_t)(a_lo + 2 * kg), d_up | (uint64_t)(b_lo + 2 * kg), idesc);
                    accum = 1;
                    umma_commit(&empty_bar[s]);  // frees the stage when these MMAs have read it
                }
                a_lo += A_STAGE_BYTES >> 4;
                b_lo += C::B_STAGE_BYTES >> 4;
                if (++s == C::STAGES) { s = 0; ph ^= 1; a_lo = a0; b_lo = b0; }
            }
            umma_commit(tmem_full_bar);
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..5 own the TMEM lane quarter (warp % 4); one thread = one output pixel =====
        const int q = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int mrow = q * 32 + lane;
        const int oy = y0 + (mrow >> p.tw_log2), ox = x0 + (mrow & (TW - 1));
        const bool pix_ok = (oy < p.Ho) && (ox < p.Wo);
        float* ypix = p.y + (long long)b * p.y_sB + (long long)oy * p.y_sH + (long long)ox * p.y_sW;
        const bool vec_ok = ((p.Cout & 3) == 0) && ((reinterpret_cast<uintptr_t>(ypix) & 15) == 0);
        constexpr int CH = (N_TILE >= 32) ? 32 : 16;
#pragma unroll 1
        for (int c0 = 0; c0 < N_TILE; c0 += CH) {
            uint32_t r[CH];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            if constexpr (CH == 32) tmem_ld32(taddr, r);
            else tmem_ld16(taddr, r);
            tmem_ld_wait();
            float v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
            if (n_issuers == 2) {   // (uniform) the odd iterations' accumulator
                if constexpr (CH == 32) tmem_ld32(taddr + (uint32_t)N_TILE, r);
                else tmem_ld16(taddr + (uint32_t)N_TILE, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CH; ++j) v[j] += __uint_as_float(r[j]);
            }
            if (!pix_ok) continue;
            bias_act<CH>(v, p.bias, n0 + c0, p.Cout, p.act, p.slope);
            if (vec_ok) {
#pragma unroll
                for (int j = 0; j < CH; j += 4)
                    if (n0 + c0 + j < p.Cout)
                        *reinterpret_cast<float4*>(ypix + n0 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    if (n0 + c0 + j < p.Cout) ypix[n0 + c0 + j] = v[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
        const float* src = reinterpret_cast<const float*>(smA);
        for (int i = threadIdx.x; i < A_STAGE_BYTES / 4; i += IGEMM_THREADS) p.dbg[i] = src[i];
        src = reinterpret_cast<const float*>(smB);
        for (int i = threadIdx.x; i < C::B_STAGE_BYTES / 4; i += IGEMM_THREADS) p.dbg[A_STAGE_BYTES / 4 + i] = src[i];
        if (threadIdx.x == 0) p.dbg[A_STAGE_BYTES / 4 + C::B_STAGE_BYTES / 4] = __uint_as_float(tmem_d);
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, C::TMEM_COLS);
    }
}


// ---- data gradient of a stride-2 convolution (default since round 2; MVF_DGRAD_S2=0 routes to cuDNN for A/B) ---
// gx[iy, ix, ci] = sum over (kh, kw, co) with iy = 2*oy + kh - pad, ix = 2*ox + kw - pad of gy[oy, ox, co] * w[co, ci, kh, kw].
// The input pixels fall into four parity classes (py, px) = (iy & 1, ix & 1); within a class only the taps with
// kh = py + pad (mod 2), kw = px + pad (mod 2) contribute and pixel (2i + py, 2j + px) reads gy at (i + dy, j + dx) with
// dy = (py + pad - kh) / 2: every class is a small stride-1 convolution over gy (<= 4 taps for 3x3, 1 for 1x1) whose
// output lands on a stride-2 lattice of gx.  One persistent launch walks the tiles of all non-empty classes with the
// skeleton of conv_patch_kernel (persistent tiles, two accumulator buffers in TMEM); gy boxes that leave the image are zero-filled by TMA.  Classes without taps
// (1x1 stride 2: three of four) are not visited -- the caller provides a zeroed gx in that case.
struct S2Class {
    int py, px, ntaps;
    int dy[16], dx[16], tap[16];  // tap = index into the dgrad-packed (flipped) bank: (KH-1-kh) * KW + (KW-1-kw)
};
struct S2Plan {
    int n;  // non-empty classes
    S2Class c[4];
};

S2Plan plan_dgrad_s2(int KH, int KW, int pad) {
    S2Plan pl = {};
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            S2Class c = {};
            c.py = py;
            c.px = px;
            for (int kh = 0; kh < KH; ++kh) {
                if ((py + pad - kh) & 1) continue;
                for (int kw = 0; kw < KW; ++kw) {
                    if ((px + pad - kw) & 1) continue;
                    if (c.ntaps >= 16) continue;  // callers check conv_dgrad_s2_check first
                    c.dy[c.ntaps] = (py + pad - kh) / 2;  // exact: the numerator is even
                    c.dx[c.ntaps] = (px + pad - kw) / 2;
                    c.tap[c.ntaps] = (KH - 1 - kh) * KW + (KW - 1 - kw);
                    ++c.ntaps;
                }
            }
            if (c.ntaps > 0) pl.c[pl.n++] = c;
        }
    return pl;
}

template <int N_TILE>
__global__ void __launch_bounds__(NTHREADS) conv_dgrad_s2_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB, const ConvArgs p,
                                                                 const __grid_constant__ S2Plan plan, const int n_mtiles,
                                                                 const int n_ntiles) {
    using C = Cfg<N_TILE>;
    constexpr int TMEM_COLS = (2 * N_TILE) < 32 ? 32 : (2 * N_TILE);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* smA = smem;
    unsigned char* smB = smem + C::STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smB + C::STAGES * C::B_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* acc_full = empty_bar + C::STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 1 << p.tw_log2, TH = TILE_M >> p.tw_log2;
    const int per_class = n_mtiles * n_ntiles, n_tiles = per_class * plan.n;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int ci = t / per_class, rem = t - ci * per_class;
                const S2Class& cl = plan.c[ci];
                const int m = rem % n_mtiles, nt = rem / n_mtiles;
                const int tx = m % p.tiles_x, ty = (m / p.tiles_x) % p.tiles_y, b = m / (p.tiles_x * p.tiles_y);
                const int x0 = tx * TW, y0 = ty * TH, n0 = nt * N_TILE;
                for (int tp = 0; tp < cl.ntaps; ++tp)
                    for (int cb = 0; cb < p.n_cblk; ++cb) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + C::B_STAGE_BYTES);
                        // A: gy as (co, ox, oy, b), box (32, TW, TH, 1) at the class's tile origin shifted by the tap
                        tma_load_4d(smA + s * A_STAGE_BYTES, &mapA, &full_bar[s], cb * BLOCK_K, x0 + cl.dx[tp], y0 + cl.dy[tp], b);
                        tma_load_3d(smB + s * C::B_STAGE_BYTES, &mapB, &full_bar[s], 0, n0, cl.tap[tp] * p.n_cblk + cb);
                        if (++s == C::STAGES) { s = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_tf32(TILE_M, N_TILE, 0, 0);
        const uint32_t smA_u = smem_u32(smA), smB_u = smem_u32(smB);
        int s = 0, ph = 0, acc = 0, accph = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int n_iters = plan.c[t / per_class].ntaps * p.n_cblk;
            mbar_wait(&acc_empty[acc], accph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_d + (uint32_t)(acc * N_TILE);
            for (int it = 0; it < n_iters; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_base = smA_u + (uint32_t)(s * A_STAGE_BYTES), b_base = smB_u + (uint32_t)(s * C::B_STAGE_BYTES);
#pragma unroll
                    for (int kg = 0; kg < KGROUPS; ++kg) {
                        const uint64_t adesc = make_smem_desc(a_base + kg * 32, 16, 1024, SWZ_128B);
                        const uint64_t bdesc = make_smem_desc(b_base + kg * 32, 16, 1024, SWZ_128B);
                        umma_tf32(d_tmem, adesc, bdesc, idesc, (it > 0 || kg > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                }
                __syncwarp();
                if (++s == C::STAGES) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&acc_full[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; accph ^= 1; }
        }
    } else {
        const int q = warp & 3;
        const int mrow = q * 32 + lane;
        constexpr int CH = (N_TILE >= 32) ? 32 : 16;
        int acc = 0, accph = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int ci = t / per_class, rem = t - ci * per_class;
            const int m = rem % n_mtiles, nt = rem / n_mtiles;
            const int tx = m % p.tiles_x, ty = (m / p.tiles_x) % p.tiles_y, b = m / (p.tiles_x * p.tiles_y);
            const int n0 = nt * N_TILE;
            // tile pixel (i, j) of the class lattice -> input pixel (2i + py, 2j + px)
            const int iy = 2 * (ty * TH + (mrow >> p.tw_log2)) + plan.c[ci].py, ix = 2 * (tx * TW + (mrow & (TW - 1))) + plan.c[ci].px;
            const bool pix_ok = (iy < p.Ho) && (ix < p.Wo);
            float* ypix = p.y + (long long)b * p.y_sB + (long long)iy * p.y_sH + (long long)ix * p.y_sW;
            const bool vec_ok = ((p.Cout & 3) == 0) && ((reinterpret_cast<uintptr_t>(ypix) & 15) == 0);
            mbar_wait(&acc_full[acc], accph);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < N_TILE; c0 += CH) {
                uint32_t r[CH];
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N_TILE + c0);
                if constexpr (CH == 32) tmem_ld32(taddr, r);
                else tmem_ld16(taddr, r);
                tmem_ld_wait();
                if (!pix_ok || n0 + c0 >= p.Cout) continue;
                float v[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
                bias_act<CH>(v, p.bias, n0 + c0, p.Cout, p.act, p.slope);   // transposed-convolution use: bias of nn.ConvTranspose2d
                if (vec_ok) {
#pragma unroll
                    for (int j = 0; j < CH; j += 4)
                        if (n0 + c0 + j < p.Cout) *reinterpret_cast<float4*>(ypix + n0 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (n0 + c0 + j < p.Cout) ypix[n0 + c0 + j] = v[j];
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; accph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TMEM_COLS);
    }
}


// ---- stride-1 multi-tap convolutions: one haloed input patch per 32-channel block serves all taps -----------------
// The per-tap kernel above pulls a fresh 16 KB A tile through L2 for every tap; here the producer loads, once per
// 32-channel block, the patch of R input rows x P input columns that all KH*KW taps of the CTA's outputs read
// (R = rows + KH - 1, P = columns + KW - 1).  The CTA's M rows are 128 CONSECUTIVE positions of that patch taken in
// row-major order with pitch P ("flat-pitch" tile): output position m = r*P + j reads, for tap (kh,kw), patch row
// m + kh*P + kw -- a uniform shift, so a tap is the same K-major operand with its start address moved by whole
// 128-byte rows (the tensor core evaluates the 128B swizzle on absolute shared-memory address bits, so a start that
// is not 1024-byte aligned is fine: mvf_selftest_umma_rows).  Positions with j > P - KW are padding columns whose
// accumulator rows are never stored.  MT such 128-position tiles (stacked vertically, sharing one patch and every
// filter tile) are accumulated side by side in TMEM to halve the filter traffic per MAC.
struct PatchArgs {
    float* y;
    const float* bias;
    const float* slope;          // act == 3: per-channel PReLU slopes
    long long y_sB, y_sH, y_sW;
    int Cout, Ho, Wo;
    int KH, KW, pad;
    int P, PWo, TR, R;           // patch pitch, valid output columns per row, output rows per 128-position tile, patch rows
    int nseg, tiles_y;           // column segments per image row, CTA rows per image
    int n_cblk, act;
    int nb;                      // filter-ring slots in use (<= NB): as many as shared memory allows, TMA latency is what they hide
    int kg_last;                 // 8-channel MMA groups that hold real channels in the LAST 32-channel block (1..4)
    int patch_bytes, patch_stride;  // bytes landed per patch, distance between the two patch buffers (1024-aligned)
    int n_mtiles, n_tiles;       // M tiles (B * nseg * tiles_y), all tiles (M tiles x N tiles)
    int b_resident;              // 1: the whole filter bank of the CTA's couts fits in the ring and is loaded ONCE per CTA
    int tma_store;               // 1: epilogue stages slabs in shared memory and stores them with TMA; 0: per-thread stores
    int no_split;                // MT = 1: 1 = issuer 0 multiplies every tap (MVF_PATCH_SPLIT=0, A/B timing and bisection)
    int dbg;                     // timing experiments only: 4 = no MMAs, 8 = no TMA loads
};

// Persistent: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA producer runs ahead across
// tile boundaries (patch double buffer + NB-deep filter ring), the MMA issuer alternates between two accumulator
// buffers in TMEM, and the four epilogue warps drain one buffer while the next tile is being multiplied.
template <int N_TILE, int MT, int NB>
__global__ void __launch_bounds__(PATCH_THREADS, 1) conv_patch_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB,
                                                                 const __grid_constant__ CUtensorMap mapY, const PatchArgs p) {
    constexpr int B_STAGE_BYTES = N_TILE * BLOCK_K * 4;
    constexpr int ACC_COLS = 2 * N_TILE;   // one accumulator buffer: MT = 2: the two stacked tiles; MT = 1: the even / odd taps' partial sums
    constexpr int TMEM_COLS = (2 * ACC_COLS) < 32 ? 32 : (2 * ACC_COLS);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int SLAB = N_TILE < 32 ? N_TILE : 32;      // output channels per staged slab (one TMA-store box row = SLAB floats)
    constexpr int STAGE_OUT_BYTES = TILE_M * SLAB * 4;   // 16 KB (8 KB for 16-channel tiles)
    unsigned char* smB = smem;
    unsigned char* smO = smem + p.nb * B_STAGE_BYTES;    // two output staging buffers (p.nb <= NB ring slots in use)
    unsigned char* smA = smO + 2 * STAGE_OUT_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smA + 2 * p.patch_stride);
    uint64_t* a_empty = a_full + 2;
    uint64_t* b_full = a_empty + 2;
    uint64_t* b_empty = b_full + NB;
    uint64_t* acc_full = b_empty + NB;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = p.KH * p.KW;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 2);      // one commit per MMA issuer
            mbar_init(&acc_full[s], 2);
            mbar_init(&acc_empty[s], 128);  // every epilogue thread arrives
        }
        for (int s = 0; s < NB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 2);      // both issuers release a filter tile (MT = 1: the one whose tap it is not, right after seeing it)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    pdl_sync();  // prologue done: wait for the producer of our inputs, let the next kernel start its own prologue

    if (warp == 0) {
        // ===== TMA producer (one lane) =====
        if (lane == 0) {
            int bs = 0, bph = 0;  // filter ring slot / phase
            int ab = 0, aph = 0;  // patch buffer / phase
            if (p.b_resident) {
                // small layers: all (tap, channel-block) filter tiles stay in shared memory for the CTA's lifetime
                const int nb_tiles = taps * p.n_cblk;
                mbar_arrive_expect_tx(&b_full[0], nb_tiles * B_STAGE_BYTES);
                for (int i = 0; i < nb_tiles; ++i) tma_load_3d(smB + i * B_STAGE_BYTES, &mapB, &b_full[0], 0, 0, i);
            }
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                const int m = t % p.n_mtiles, nt = t / p.n_mtiles;
                const int seg = m % p.nseg, ty = (m / p.nseg) % p.tiles_y, b = m / (p.nseg * p.tiles_y);
                const int xs = seg * p.PWo, ys = ty * (MT * p.TR), n0 = nt * N_TILE;
                for (int cb = 0; cb < p.n_cblk; ++cb) {
                    mbar_wait(&a_empty[ab], aph ^ 1);
                    if (p.dbg & 8) {
                        mbar_arrive(&a_full[ab]);
                    } else {
                        mbar_arrive_expect_tx(&a_full[ab], p.patch_bytes);
                        // patch: dims (c, x, y, b), box (32, P, R, 1) at the tile's input origin -> smem [R*P pixels][32 c]
                        tma_load_4d(smA + ab * p.patch_stride, &mapA, &a_full[ab], cb * BLOCK_K, xs - p.pad, ys - p.pad, b);
                    }
                    if (++ab == 2) { ab = 0; aph ^= 1; }
                    if (p.b_resident) continue;
                    int kb = cb;  // filter tile index = tap * n_cblk + cb
                    for (int tp = 0; tp < taps; ++tp, kb += p.n_cblk) {
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        if (p.dbg & 8) {
                            mbar_arrive(&b_full[bs]);
                        } else {
                            mbar_arrive_expect_tx(&b_full[bs], B_STAGE_BYTES);
                            tma_load_3d(smB + bs * B_STAGE_BYTES, &mapB, &b_full[bs], 0, n0, kb);
                        }
                        if (++bs == p.nb) { bs = 0; bph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 6) {
        // ===== MMA issuers: warp 1 owns the stacked tile mt = 0, warp 6 the tile mt = 1; ONE lane of each runs the whole loop =====
        // Measured (tools/conv_sweep.py, MVF_CONV_DBG=8: no TMA loads at all): every layer shape ran exactly as long without its
        // loads as with them, and ncu put the tensor pipe at 40 % of the active cycles with the operand fetch at its ideal 64
        // wavefronts per MMA -- the kernel is bound by the ISSUE of the MMAs: one thread sustains one tcgen05.mma per 60-80 cycles
        // (N = 64: 32 cycles of tensor work, N = 128: 64) plus ~220 cycles of barrier / descriptor bookkeeping per tap.  So the
        // issue work is (a) thinned -- descriptor words advance by adds, the accumulate flag is a register that flips once per
        // tile, full 32-channel blocks take an unpredicated path, no ELECT / VOTE / BRA.DIV per tap -- and (b) split: each stacked
        // tile has its own accumulator columns and its own issuing warp, so two instruction streams feed the tensor pipe and the
        // order of the additions into any one accumulator stays fixed (results are bitwise repeatable).  The barriers the
        // issuers release (a_empty, b_empty, acc_full) count MT arrivals.
        // With a single tile per CTA (MT = 1) the two issuers take alternate taps (counted across channel blocks) into two partial
        // accumulators that the epilogue adds.
        const int my_mt = (warp == 1) ? 0 : 1;
        constexpr bool SPLIT_TAPS = (MT == 1);
        if (elect_one()) {   // (elect.sync rather than lane == 0: the compiler then knows a single lane is active and moves the operands to
                             //  uniform registers with one R2UR each instead of a per-MMA ELECT / R2UR.BROADCAST loop)
            constexpr uint32_t idesc = make_idesc_tf32(TILE_M, N_TILE, 0, 0);
            const uint64_t desc_hi = make_smem_desc(0, 16, 1024, SWZ_128B);  // everything but the start address
            const uint64_t d_up = desc_hi & 0xFFFFFFFF00000000ull;
            const uint32_t lo_const = (uint32_t)desc_hi;                     // LBO field; the start address (>> 4) is OR-ed / added below
            const uint32_t smA_u = smem_u32(smA), smB_u = smem_u32(smB);
            const uint32_t mt_step = SPLIT_TAPS ? 0u : (uint32_t)(p.TR * p.P) * 8u;  // 16-byte units between the stacked tiles' first rows
            const uint32_t row_step = (uint32_t)(p.P - p.KW) * 8u;           // extra advance at the end of a filter row
            constexpr uint32_t B_UNITS = B_STAGE_BYTES >> 4;
            const uint32_t b_ring0 = (smB_u >> 4) | lo_const;
            const int resident = p.b_resident;
            int bs = 0, bph = 0, ab = 0, aph = 0, acc = 0, accph = 0;
            uint32_t b_ring = b_ring0;
            if (resident) {
                mbar_wait(&b_full[0], 0);
                tc_fence_after();
            }
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                mbar_wait(&acc_empty[acc], accph ^ 1);  // the epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t d_base = tmem_d + (uint32_t)(acc * ACC_COLS + my_mt * N_TILE);
                uint32_t accum = 0;
                int par = 0;                            // parity of the tap counter within the tile
                for (int cb = 0; cb < p.n_cblk; ++cb) {
                    mbar_wait(&a_full[ab], aph);
                    tc_fence_after();
                    uint32_t a_tap = (((smA_u + (uint32_t)(ab * p.patch_stride)) >> 4) | lo_const) + (uint32_t)my_mt * mt_step;
                    uint32_t b_res = b_ring0 + (uint32_t)cb * B_UNITS;        // resident bank: tile (tap, cb) at (tap * n_cblk + cb)
                    const uint32_t b_res_step = (uint32_t)p.n_cblk * B_UNITS;
                    const int ng = (cb == p.n_cblk - 1) ? p.kg_last : KGROUPS;  // all-zero channel groups are not multiplied
                    int kw = 0;
                    for (int tp = 0; tp < taps; ++tp) {
                        const bool mine = !SPLIT_TAPS || (p.no_split ? my_mt == 0 : par == my_mt);
                        par ^= 1;
                        uint32_t b_lo;
                        if (resident) {
                            b_lo = b_res;
                            b_res += b_res_step;
                        } else {
                            // both issuers wait on EVERY filter-tile barrier, in order (a thread that skipped a phase of a barrier
                            // could later find it one phase behind and take that for "complete": TMA completions are not ordered)
                            mbar_wait(&b_full[bs], bph);
                            if (mine) tc_fence_after();
                            b_lo = b_ring;
                        }
                        if (mine && !(p.dbg & 4)) {
                            umma_tf32(d_base, d_up | (uint64_t)a_tap, d_up | (uint64_t)b_lo, idesc, accum);
                            if (ng == KGROUPS) {
#pragma unroll
                                for (int kg = 1; kg < KGROUPS; ++kg)
                                    umma_tf32_acc(d_base, d_up | (uint64_t)(a_tap + 2 * kg), d_up | (uint64_t)(b_lo + 2 * kg), idesc);
                            } else {
#pragma unroll
                                for (int kg = 1; kg < KGROUPS; ++kg)
                                    if (kg < ng) umma_tf32_acc(d_base, d_up | (uint64_t)(a_tap + 2 * kg), d_up | (uint64_t)(b_lo + 2 * kg), idesc);
                            }
                        }
                        if (mine) accum = 1;
                        if (!resident) {
                            // (MT = 1) the other issuer only reports that it has seen this phase: try_wait.parity tells the current
                            // phase from the preceding one only, so no issuer may fall a ring generation behind
                            if (mine) umma_commit(&b_empty[bs]);
                            else mbar_arrive(&b_empty[bs]);
                            b_ring += B_UNITS;
                            if (++bs == p.nb) { bs = 0; bph ^= 1; b_ring = b_ring0; }
                        }
                        a_tap += 8;
                        if (++kw == p.KW) { kw = 0; a_tap += row_step; }
                    }
                    umma_commit(&a_empty[ab]);
                    if (++ab == 2) { ab = 0; aph ^= 1; }
                }
                umma_commit(&acc_full[acc]);
                if (++acc == 2) { acc = 0; accph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..5 own the TMEM lane quarter (warp % 4); one thread = one patch position =====
        // TMEM -> registers -> bias / activation -> swizzled shared-memory slab [128 positions][SLAB channels] -> one TMA
        // store per output row of the tile (the TMA unit clips rows / columns outside the image and the padding columns
        // are simply not part of the box), so global memory sees full 128-byte lines instead of per-thread fragments.
        const int q = warp & 3;
        const int mrow = q * 32 + lane;
        const int et = threadIdx.x - 64;  // 0..127 among the epilogue threads
        const int r = mrow / p.P, j = mrow - r * p.P;
        const bool use_tma = p.tma_store != 0;
        int acc = 0, accph = 0, sb = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            const int m = t % p.n_mtiles, nt = t / p.n_mtiles;
            const int seg = m % p.nseg, ty = (m / p.nseg) % p.tiles_y, b = m / (p.nseg * p.tiles_y);
            const int xs = seg * p.PWo, ys = ty * (MT * p.TR), n0 = nt * N_TILE;
            mbar_wait(&acc_full[acc], accph);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int oy = ys + mt * p.TR + r, ox = xs + j;
                const bool pix_ok = (r < p.TR) && (j < p.PWo) && (oy < p.Ho) && (ox < p.Wo);
                float* ypix = p.y + (long long)b * p.y_sB + (long long)oy * p.y_sH + (long long)ox * p.y_sW;
#pragma unroll 1
                for (int c0 = 0; c0 < N_TILE; c0 += SLAB) {
                    uint32_t rr[SLAB];
                    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + mt * N_TILE + c0);
                    if constexpr (SLAB == 32) tmem_ld32(taddr, rr);
                    else tmem_ld16(taddr, rr);
                    tmem_ld_wait();
                    if (n0 + c0 >= p.Cout) continue;  // (uniform) slab entirely past Cout
                    float v[SLAB];
#pragma unroll
                    for (int jj = 0; jj < SLAB; ++jj) v[jj] = __uint_as_float(rr[jj]);
                    if (MT == 1 && !p.no_split) {     // the odd taps' partial sums
                        if constexpr (SLAB == 32) tmem_ld32(taddr + (uint32_t)N_TILE, rr);
                        else tmem_ld16(taddr + (uint32_t)N_TILE, rr);
                        tmem_ld_wait();
#pragma unroll
                        for (int jj = 0; jj < SLAB; ++jj) v[jj] += __uint_as_float(rr[jj]);
                    }
                    bias_act<SLAB>(v, p.bias, n0 + c0, p.Cout, p.act, p.slope);
                    if (use_tma) {
                        // the staging buffer must have been read by the TMA store issued two slabs ago
                        if (et == 0) tma_store_wait_read<1>();
                        named_barrier_sync(1, 128);
                        unsigned char* row = smO + sb * STAGE_OUT_BYTES + mrow * (SLAB * 4);
#pragma unroll
                        for (int k = 0; k < SLAB / 4; ++k) {
                            // 128B (64B) swizzle on absolute address bits, as the TMA store expects
                            const int kk = (SLAB == 32) ? (k ^ (mrow & 7)) : (k ^ ((mrow >> 1) & 3));
                            *reinterpret_cast<float4*>(row + kk * 16) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        }
                        fence_proxy_async();
                        named_barrier_sync(1, 128);
                        if (et == 0) {
                            for (int rrow = 0; rrow < p.TR; ++rrow) {
                                const int oyr = ys + mt * p.TR + rrow;
                                if (oyr < p.Ho)
                                    tma_store_4d(&mapY, smO + sb * STAGE_OUT_BYTES + rrow * p.P * (SLAB * 4), n0 + c0, xs, oyr, b);
                            }
                            tma_store_commit();
                        }
                        sb ^= 1;
                    } else if (pix_ok) {
                        if (((p.Cout & 3) == 0) && ((reinterpret_cast<uintptr_t>(ypix) & 15) == 0)) {
#pragma unroll
                            for (int jj = 0; jj < SLAB; jj += 4)
                                if (n0 + c0 + jj < p.Cout)
                                    *reinterpret_cast<float4*>(ypix + n0 + c0 + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
                        } else {
#pragma unroll
                            for (int jj = 0; jj < SLAB; ++jj)
                                if (n0 + c0 + jj < p.Cout) ypix[n0 + c0 + jj] = v[jj];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);  // this thread's TMEM reads of the buffer are complete
            if (++acc == 2) { acc = 0; accph ^= 1; }
        }
        if (use_tma && et == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TMEM_COLS);
    }
}

// ---- filter-bank packing -------------------------------------------------------------------------------------
// fprop: Wp[tap][cb][co][kk] = w[co][cb*32+kk][kh][kw]                      (N = Cout, K = Cin)
// dgrad: Wp[tap][cb][ci][kk] = w[cb*32+kk][ci][KH-1-kh][KW-1-kw]            (N = Cin,  K = Cout)
__global__ void pack_filters_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KH, int KW,
                                    int dgrad) {
    pdl_sync();
    const int N = dgrad ? Cin : Cout, K = dgrad ? Cout : Cin;
    const int ncb = (K + BLOCK_K - 1) / BLOCK_K;
    const long long total = (long long)KH * KW * ncb * N * BLOCK_K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % BLOCK_K);
        long long r = i / BLOCK_K;
        const int n = (int)(r % N);
        r /= N;
        const int cb = (int)(r % ncb);
        const int tap = (int)(r / ncb);
        const int k = cb * BLOCK_K + kk;
        int kh = tap / KW, kw = tap - kh * KW;
        float v = 0.f;
        if (k < K) {
            if (dgrad) {
                kh = KH - 1 - kh;
                kw = KW - 1 - kw;
                v = w[(((long long)k * Cin + n) * KH + kh) * KW + kw];
            } else {
                v = w[(((long long)n * Cin + k) * KH + kh) * KW + kw];
            }
        }
        out[i] = v;
    }
}

// every filter bank of a step in ONE launch: table[e] = {w, out, Cout, Cin, KH, KW, dgrad, first block}; a CTA packs PACK_CHUNK
// consecutive elements of one bank (entry found by binary search over the first-block column)
constexpr int PACK_CHUNK = 2048;
__global__ void __launch_bounds__(256) pack_filters_multi_kernel(const long long* __restrict__ table, int n_entries) {
    int lo = 0, hi = n_entries - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid * 8 + 7] <= (long long)blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    const long long* e = table + lo * 8;
    const float* __restrict__ w = reinterpret_cast<const float*>(e[0]);
    float* __restrict__ out = reinterpret_cast<float*>(e[1]);
    const int Cout = (int)e[2], Cin = (int)e[3], KH = (int)e[4], KW = (int)e[5], dgrad = (int)e[6];
    const int N = dgrad ? Cin : Cout, K = dgrad ? Cout : Cin;
    const int ncb = (K + BLOCK_K - 1) / BLOCK_K;
    const long long total = (long long)KH * KW * ncb * N * BLOCK_K;
    const long long i0 = ((long long)blockIdx.x - e[7]) * PACK_CHUNK;
    for (long long i = i0 + threadIdx.x; i < min(total, i0 + PACK_CHUNK); i += 256) {
        const int kk = (int)(i % BLOCK_K);
        long long r = i / BLOCK_K;
        const int n = (int)(r % N);
        r /= N;
        const int cb = (int)(r % ncb);
        const int tap = (int)(r / ncb);
        const int k = cb * BLOCK_K + kk;
        int kh = tap / KW, kw = tap - kh * KW;
        float v = 0.f;
        if (k < K) {
            if (dgrad) {
                kh = KH - 1 - kh;
                kw = KW - 1 - kw;
                v = w[(((long long)k * Cin + n) * KH + kh) * KW + kw];
            } else {
                v = w[(((long long)n * Cin + k) * KH + kh) * KW + kw];
            }
        }
        out[i] = v;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

template <int N_TILE>
cudaError_t launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const ConvArgs& a, int B, cudaStream_t st) {
    using C = Cfg<N_TILE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(B * a.tiles_x * a.tiles_y, (a.Cout + N_TILE - 1) / N_TILE);
    return launch_pdl(conv_igemm_kernel<N_TILE>, grid, dim3(IGEMM_THREADS), C::SMEM_BYTES, st, mapA, mapB, a);
}

template <int N_TILE, int MT>
cudaError_t launch_patch(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapY, const PatchArgs& a,
                         cudaStream_t st) {
    constexpr int NB = (N_TILE >= 128) ? 8 : 16;  // ring slots (barriers); a.nb of them are backed by shared memory
    const int smem = a.nb * N_TILE * BLOCK_K * 4 + 2 * TILE_M * (N_TILE < 32 ? N_TILE : 32) * 4 + 2 * a.patch_stride + 1024 + 256;
    static int attr_max = 0;
    if (smem > attr_max) {
        cudaError_t e = cudaFuncSetAttribute(conv_patch_kernel<N_TILE, MT, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_max = smem;
    }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    const int grid = a.n_tiles < n_sm ? a.n_tiles : n_sm;
    return launch_pdl(conv_patch_kernel<N_TILE, MT, NB>, dim3(grid), dim3(PATCH_THREADS), (size_t)smem, st, mapA, mapB, mapY, a);
}

}  // namespace

size_t packed_filter_floats(int N, int K, int KH, int KW) {
    return (size_t)KH * KW * ((K + BLOCK_K - 1) / BLOCK_K) * N * BLOCK_K;
}

cudaError_t pack_filters(const float* w, float* out, int Cout, int Cin, int KH, int KW, int dgrad, cudaStream_t st) {
    const size_t total = packed_filter_floats(dgrad ? Cin : Cout, dgrad ? Cout : Cin, KH, KW);
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads < 1184 ? (total + threads - 1) / threads : 1184);
    launch_pdl(pack_filters_kernel, dim3(blocks), dim3(threads), 0, st, w, out, Cout, Cin, KH, KW, dgrad);
    return cudaGetLastError();
}

int pack_chunk() { return PACK_CHUNK; }
cudaError_t pack_filters_multi(const long long* table, int n_entries, long long total_blocks, cudaStream_t st) {
    if (total_blocks <= 0 || n_entries <= 0) return cudaSuccess;
    pack_filters_multi_kernel<<<(unsigned)total_blocks, 256, 0, st>>>(table, n_entries);
    return cudaGetLastError();
}

static int out_size(int n, int k, int pad, int stride) { return (n + 2 * pad - k) / stride + 1; }

const char* conv_check(const ConvDesc& d) {
    if (d.B <= 0 || d.Cin <= 0 || d.H <= 0 || d.W <= 0 || d.Cout <= 0 || d.KH <= 0 || d.KW <= 0) return "non-positive size";
    if ((d.stride != 1 && d.stride != 2) || (d.stride_x != 1 && d.stride_x != 2)) return "strides must be 1 or 2";
    if (d.Cin % 4 != 0) return "input channels must be a multiple of 4 (TMA: 16-byte rows)";
    if ((d.x_sH % 4) || (d.x_sW % 4) || (d.x_sB % 4)) return "input strides must be multiples of 4 elements (TMA: 16 bytes)";
    if (d.pad < 0) return "negative padding";
    if (d.H + 2 * d.pad < d.KH || d.W + 2 * d.pad < d.KW) return "empty output";
    return nullptr;
}

static float* g_dbg = nullptr;
void set_debug_buffer(float* p) { g_dbg = p; }

static cudaError_t conv_forward_patch(const ConvDesc& d, const float* x, const float* w_packed, const float* bias, float* y,
                                      int act, cudaStream_t st, const char** why, EncodeTiledFn enc, int Ho, int Wo, const float* slope) {
    // column segmentation: the split of an output row into nseg pieces that needs the fewest 128-position tiles
    int best_nseg = 1;
    long long best_tiles = -1;
    for (int nseg = 1; nseg <= 16; ++nseg) {
        const int pwo = (Wo + nseg - 1) / nseg, P = pwo + d.KW - 1;
        if (P > TILE_M) continue;
        const int TR = TILE_M / P;
        const long long tiles = (long long)nseg * ((Ho + TR - 1) / TR);
        if (best_tiles < 0 || tiles < best_tiles) {
            best_tiles = tiles;
            best_nseg = nseg;
        }
    }
    if (best_tiles < 0) {
        *why = "output row too wide for the patch kernel";
        return cudaErrorInvalidValue;
    }
    PatchArgs a;
    a.y = y; a.bias = bias; a.slope = slope;
    a.y_sB = d.y_sB; a.y_sH = d.y_sH; a.y_sW = d.y_sW;
    a.Cout = d.Cout; a.Ho = Ho; a.Wo = Wo; a.KH = d.KH; a.KW = d.KW; a.pad = d.pad;
    a.nseg = best_nseg;
    a.PWo = (Wo + best_nseg - 1) / best_nseg;
    a.P = a.PWo + d.KW - 1;
    a.TR = TILE_M / a.P;
    // Tile shape (N_TILE output channels x MT stacked 128-position tiles per CTA) from a cost model.  Measured on B200
    // (profiles/r2_conv_layers_cfg2.txt): the C >= 64 layers run at the L2 -> SM bandwidth of the bytes their tiles pull in -- every
    // tile re-reads its (taps x Cin x N_TILE) filter slice, 4-8x the patch bytes -- about 5 TB/s aggregate, not at the MMA rate:
    //   64->64 @48x160: 114 MB / 30.6 us, 128->128 @24x80: 146 MB / 31.3 us, 256->256 @12x40: 140 MB / 31.1 us, 512->512 @6x20: 263 MB / 51.3 us.
    // The model charges a candidate max(bytes / L2 rate, waves x max(MMA cycles, per-SM ingest cycles)) and takes the cheapest; the
    // round-1 rule (widest N <= 128, MT = 2 only with >= 2 waves of CTAs) remains as MVF_CONV_TILE_RULE=r1 for A/B timing.
    int n_tile = 16, MT = 1;
    {
        int n_wide = 16;
        while (n_wide < d.Cout && n_wide < 128) n_wide *= 2;
        const int taps = d.KH * d.KW, n_cblk = (d.Cin + BLOCK_K - 1) / BLOCK_K;
        const char* rule = getenv("MVF_CONV_TILE_RULE");
        if (rule && rule[0] == 'r') {
            n_tile = n_wide;
            const long long ctas2 = (long long)d.B * best_nseg * ((Ho + 2 * a.TR - 1) / (2 * a.TR)) * ((d.Cout + n_tile - 1) / n_tile);
            if (ctas2 >= 2 * 148 && Ho >= 2 * a.TR) MT = 2;
        } else if (rule) {  // "N,MT": forced shape (tools/conv_layers.py sweeps)
            n_tile = atoi(rule);
            const char* c = strchr(rule, ',');
            MT = c ? atoi(c + 1) : 1;
            if ((n_tile != 16 && n_tile != 32 && n_tile != 64 && n_tile != 128) || (MT != 1 && MT != 2)) {
                *why = "MVF_CONV_TILE_RULE: N in {16,32,64,128}, MT in {1,2}";
                return cudaErrorInvalidValue;
            }
        } else {
            double best_cost = -1.0;
            for (int nt = 16; nt <= n_wide; nt *= 2)
                for (int mt = 1; mt <= 2; ++mt) {
                    if (mt == 2 && Ho < 2 * a.TR) continue;
                    const long long R = mt * a.TR + d.KH - 1;
                    const long long patch = R * a.P * BLOCK_K * 4;
                    const long long pstride = ((patch + (d.KW - 1 + TILE_M - a.TR * a.P) * 128) + 1023) / 1024 * 1024;
                    const int slab = nt < 32 ? nt : 32;
                    long long ring = (227 * 1024 - 1280 - 2 * TILE_M * slab * 4 - 2 * pstride) / (nt * BLOCK_K * 4);
                    if (ring < 2) continue;
                    if (ring > (nt >= 128 ? 8 : 16)) ring = nt >= 128 ? 8 : 16;
                    const long long n_nt = (d.Cout + nt - 1) / nt;
                    const long long tiles = (long long)d.B * best_nseg * ((Ho + mt * a.TR - 1) / (mt * a.TR)) * n_nt;
                    const long long ctas = tiles < 148 ? tiles : 148;
                    const long long per_cta = (tiles + ctas - 1) / ctas;
                    const bool resident = n_nt == 1 && taps * n_cblk <= ring;
                    const double filt = (double)taps * n_cblk * nt * BLOCK_K * 4;
                    const double a_bytes = (double)n_cblk * patch, o_bytes = (double)mt * TILE_M * nt * 4;
                    const double l2_bytes = tiles * (a_bytes + o_bytes + (resident ? 0.0 : filt)) + (resident ? ctas * filt : 0.0);
                    // per tap: each issuer warp (one per stacked tile) spends ~220 cycles of bookkeeping + ~75 per MMA; the tensor pipe
                    // needs nt / 2 cycles per MMA of all stacked tiles
                    const double issue_one = 220.0 + KGROUPS * (nt / 2 > 75 ? nt / 2 : 75.0), exec_tap = (double)mt * KGROUPS * (nt / 2);
                    const double issue_tap = mt == 1 ? 0.5 * issue_one + 60.0 : issue_one;   // MT = 1: the issuers take alternate taps
                    const double mma_clk = (double)taps * n_cblk * (issue_tap > exec_tap ? issue_tap : exec_tap);
                    const double ingest_clk = (a_bytes + (resident ? 0.0 : filt)) / 48.0;
                    const double sm_clk = per_cta * (mma_clk > ingest_clk ? mma_clk : ingest_clk) + 3000.0 * per_cta + 4000.0;
                    const double l2_clk = l2_bytes / 2500.0;
                    const double cost = sm_clk > l2_clk ? sm_clk : l2_clk;
                    if (best_cost < 0.0 || cost < best_cost) {
                        best_cost = cost;
                        n_tile = nt;
                        MT = mt;
                    }
                }
            if (best_cost < 0.0) {
                *why = "patch does not fit in shared memory";
                return cudaErrorInvalidValue;
            }
        }
    }
    const int n_ntiles = (d.Cout + n_tile - 1) / n_tile;
    a.R = MT * a.TR + d.KH - 1;
    a.tiles_y = (Ho + MT * a.TR - 1) / (MT * a.TR);
    a.n_cblk = (d.Cin + BLOCK_K - 1) / BLOCK_K;
    a.act = act;
    a.patch_bytes = a.R * a.P * BLOCK_K * 4;
    // the MMAs of the last taps read up to (KW - 1) rows past R*P for positions that are never stored; keep them inside the buffer
    a.patch_stride = ((a.patch_bytes + (d.KW - 1 + TILE_M - a.TR * a.P) * 128) + 1023) / 1024 * 1024;
    a.dbg = getenv("MVF_CONV_DBG") ? atoi(getenv("MVF_CONV_DBG")) : 0;
    a.no_split = (getenv("MVF_PATCH_SPLIT") && atoi(getenv("MVF_PATCH_SPLIT")) == 0) ? 1 : 0;
    a.n_mtiles = d.B * a.nseg * a.tiles_y;
    a.n_tiles = a.n_mtiles * n_ntiles;
    const int slab = n_tile < 32 ? n_tile : 32;
    // filter ring: up to 8 x 16 KB (N = 128) / 16 slots (narrower tiles), limited by what the patches leave free
    const int nb_max = n_tile >= 128 ? 8 : 16;
    int nb_ring = (227 * 1024 - 1280 - 2 * TILE_M * slab * 4 - 2 * a.patch_stride) / (n_tile * BLOCK_K * 4);
    if (nb_ring > nb_max) nb_ring = nb_max;
    if (const char* e = getenv("MVF_CONV_NB")) nb_ring = atoi(e) < nb_ring ? atoi(e) : nb_ring;
    a.nb = nb_ring;
    a.kg_last = ((d.Cin - (a.n_cblk - 1) * BLOCK_K) + 7) / 8;
    a.b_resident = (n_ntiles == 1 && d.KH * d.KW * a.n_cblk <= nb_ring && !getenv("MVF_CONV_NO_RESIDENT")) ? 1 : 0;
    if (nb_ring < 2) {
        *why = "patch does not fit in shared memory";
        return cudaErrorInvalidValue;
    }
    CUtensorMap mapA, mapB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.x_sW * 4, (cuuint64_t)d.x_sH * 4, (cuuint64_t)d.x_sB * 4};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)a.P, (cuuint32_t)a.R, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the input patch";
            return cudaErrorInvalidValue;
        }
    }
    {
        cuuint64_t dims[3] = {BLOCK_K, (cuuint64_t)d.Cout, (cuuint64_t)(d.KH * d.KW * a.n_cblk)};
        cuuint64_t strides[2] = {BLOCK_K * 4, (cuuint64_t)d.Cout * BLOCK_K * 4};
        cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)n_tile, 1};
        cuuint32_t es[3] = {1, 1, 1};
        if (enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(w_packed), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the packed filter bank";
            return cudaErrorInvalidValue;
        }
    }
    // output as (c, x, y, b) for the TMA-store epilogue: one box = one output row of a tile, SLAB channels wide
    CUtensorMap mapY;
    // (16-channel tiles would need 64-byte staging rows, which the TMA store cannot address at odd row offsets)
    a.tma_store = (n_tile >= 32 && (d.Cout % 4) == 0 && (d.y_sW % 4) == 0 && (d.y_sH % 4) == 0 && (d.y_sB % 4) == 0 && (((uintptr_t)y) & 15) == 0 &&
                   !getenv("MVF_CONV_NO_TMA_STORE")) ? 1 : 0;
    if (a.tma_store) {
        cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.y_sW * 4, (cuuint64_t)d.y_sH * 4, (cuuint64_t)d.y_sB * 4};
        cuuint32_t box[4] = {(cuuint32_t)slab, (cuuint32_t)a.PWo, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (enc(&mapY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                slab == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the output tensor";
            return cudaErrorInvalidValue;
        }
    } else {
        mapY = mapA;  // unused by the kernel
    }
    if (MT == 2) {
        switch (n_tile) {
            case 16: return launch_patch<16, 2>(mapA, mapB, mapY, a, st);
            case 32: return launch_patch<32, 2>(mapA, mapB, mapY, a, st);
            case 64: return launch_patch<64, 2>(mapA, mapB, mapY, a, st);
            default: return launch_patch<128, 2>(mapA, mapB, mapY, a, st);
        }
    }
    switch (n_tile) {
        case 16: return launch_patch<16, 1>(mapA, mapB, mapY, a, st);
        case 32: return launch_patch<32, 1>(mapA, mapB, mapY, a, st);
        case 64: return launch_patch<64, 1>(mapA, mapB, mapY, a, st);
        default: return launch_patch<128, 1>(mapA, mapB, mapY, a, st);
    }
}

cudaError_t conv_forward(const ConvDesc& d, const float* x, const float* w_packed, const float* bias, float* y, int act,
                         cudaStream_t st, const char** why, const float* slope) {
    *why = nullptr;
    if (act == 3 && !slope) {
        *why = "PReLU epilogue needs the slope vector";
        return cudaErrorInvalidValue;
    }
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        *why = "cuTensorMapEncodeTiled is not available from the driver";
        return cudaErrorNotSupported;
    }
    if (((uintptr_t)x & 15) || ((uintptr_t)w_packed & 15)) {
        *why = "input / packed filter pointers must be 16-byte aligned";
        return cudaErrorInvalidValue;
    }
    const int Ho = out_size(d.H, d.KH, d.pad, d.stride), Wo = out_size(d.W, d.KW, d.pad, d.stride_x);
    if (d.stride == 1 && d.stride_x == 1 && d.KH * d.KW > 1 && d.KW <= 16 && !getenv("MVF_CONV_NO_PATCH"))
        return conv_forward_patch(d, x, w_packed, bias, y, act, st, why, enc, Ho, Wo, slope);
    ConvArgs a;
    a.y = y;
    a.bias = bias;
    a.slope = slope;
    a.y_sB = d.y_sB; a.y_sH = d.y_sH; a.y_sW = d.y_sW;
    a.Cout = d.Cout; a.Ho = Ho; a.Wo = Wo;
    a.KH = d.KH; a.KW = d.KW; a.pad = d.pad; a.stride = d.stride; a.stride_x = d.stride_x;
    // tile shape: the (2^j x 128/2^j) patch that covers the output with the least padding (ties: the widest)
    int best = 5;
    long long best_cost = -1;
    for (int j = 7; j >= 1; --j) {
        const int tw = 1 << j, th = TILE_M >> j;
        const long long cost = (long long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = j;
        }
    }
    a.tw_log2 = best;
    const int TW = 1 << best, TH = TILE_M >> best;
    a.tiles_x = (Wo + TW - 1) / TW;
    a.tiles_y = (Ho + TH - 1) / TH;
    a.n_cblk = (d.Cin + BLOCK_K - 1) / BLOCK_K;
    a.act = act;
    a.dbg = g_dbg;
    a.single_issuer = (getenv("MVF_IGEMM_ISSUERS") && atoi(getenv("MVF_IGEMM_ISSUERS")) == 1) ? 1 : 0;
    int n_tile = 16;
    while (n_tile < d.Cout && n_tile < 128) n_tile *= 2;

    CUtensorMap mapA, mapB;
    {   // input as (c, x, y, b); strides in bytes for dims 1..3; the box walks x and y with the convolution stride
        cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.x_sW * 4, (cuuint64_t)d.x_sH * 4, (cuuint64_t)d.x_sB * 4};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)(TW * d.stride_x), (cuuint32_t)(TH * d.stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)d.stride_x, (cuuint32_t)d.stride, 1};
        CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the input tensor";
            return cudaErrorInvalidValue;
        }
    }
    {   // packed filter bank as (32 k, Cout, taps * n_cblk)
        cuuint64_t dims[3] = {BLOCK_K, (cuuint64_t)d.Cout, (cuuint64_t)(d.KH * d.KW * a.n_cblk)};
        cuuint64_t strides[2] = {BLOCK_K * 4, (cuuint64_t)d.Cout * BLOCK_K * 4};
        cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)n_tile, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(w_packed), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the packed filter bank";
            return cudaErrorInvalidValue;
        }
    }
    switch (n_tile) {
        case 16: return launch<16>(mapA, mapB, a, d.B, st);
        case 32: return launch<32>(mapA, mapB, a, d.B, st);
        case 64: return launch<64>(mapA, mapB, a, d.B, st);
        default: return launch<128>(mapA, mapB, a, d.B, st);
    }
}


// ---- host side of the stride-2 data gradient ------------------------------------------------------------------------
// d describes the FORWARD convolution: x_* = the gradient being produced (gx, [B, Cin, H, W]), y_* = the incoming
// gradient (gy, [B, Cout, Ho, Wo]); w_packed = mvf_conv2d_pack_filters(.., dgrad = 1) of the forward weights.
const char* conv_dgrad_s2_check(const ConvDesc& d) {
    if (d.stride != 2 || d.stride_x != 2) return "strides must be 2";
    if (d.KH > 8 || d.KW > 8) return "filter larger than 8x8";
    if (d.Cin % 4 != 0) return "Cin must be a multiple of 4 (16-byte pixels of the produced gradient)";
    if (d.Cout % 4 != 0) return "Cout must be a multiple of 4 (16-byte pixels of the incoming gradient)";
    if (d.x_sW % 4 || d.x_sH % 4 || d.x_sB % 4 || d.y_sW % 4 || d.y_sH % 4 || d.y_sB % 4) return "strides must be multiples of 4 elements";
    return nullptr;
}

int conv_dgrad_s2_plan_table(int KH, int KW, int pad, int* out, int capacity) {
    // flat copy of the plan for tests: n, then per class: py, px, ntaps, ntaps x (dy, dx, tap)
    const S2Plan pl = plan_dgrad_s2(KH, KW, pad);
    int k = 0;
    auto put = [&](int v) { if (k < capacity) out[k] = v; ++k; };
    put(pl.n);
    for (int i = 0; i < pl.n; ++i) {
        put(pl.c[i].py); put(pl.c[i].px); put(pl.c[i].ntaps);
        for (int t = 0; t < pl.c[i].ntaps; ++t) { put(pl.c[i].dy[t]); put(pl.c[i].dx[t]); put(pl.c[i].tap[t]); }
    }
    return k;
}

template <int N_TILE>
static cudaError_t launch_dgrad_s2(const CUtensorMap& mapA, const CUtensorMap& mapB, const ConvArgs& a, const S2Plan& plan, int n_mtiles,
                                   int n_ntiles, cudaStream_t st) {
    using C = Cfg<N_TILE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_dgrad_s2_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = n_mtiles * n_ntiles * plan.n;
    const int per_sm = (C::SMEM_BYTES <= 110 * 1024 && 2 * N_TILE * 2 <= 512) ? 2 : 1;
    const int ctas = n_tiles < per_sm * n_sm ? n_tiles : per_sm * n_sm;
    return launch_pdl(conv_dgrad_s2_kernel<N_TILE>, dim3(ctas), dim3(NTHREADS), C::SMEM_BYTES, st, mapA, mapB, a, plan, n_mtiles, n_ntiles);
}

cudaError_t conv_dgrad_s2(const ConvDesc& d, const float* gy, const float* w_packed, float* gx, cudaStream_t st, const char** why,
                          const float* bias) {
    *why = conv_dgrad_s2_check(d);
    if (*why) return cudaErrorInvalidValue;
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        *why = "cuTensorMapEncodeTiled is not available from the driver";
        return cudaErrorNotSupported;
    }
    if (((uintptr_t)gy & 15) || ((uintptr_t)w_packed & 15) || ((uintptr_t)gx & 15)) {
        *why = "pointers must be 16-byte aligned";
        return cudaErrorInvalidValue;
    }
    const int Ho = out_size(d.H, d.KH, d.pad, 2), Wo = out_size(d.W, d.KW, d.pad, 2);
    const S2Plan plan = plan_dgrad_s2(d.KH, d.KW, d.pad);
    if (plan.n == 0) return cudaSuccess;
    const int Hc = (d.H + 1) / 2, Wc = (d.W + 1) / 2;  // lattice of the largest class
    ConvArgs a = {};
    a.y = gx;
    a.bias = bias;   // non-null when the kernel serves as nn.ConvTranspose2d's forward (mvf_conv_transpose2d_s2_fwd)
    a.slope = nullptr;
    a.y_sB = d.x_sB; a.y_sH = d.x_sH; a.y_sW = d.x_sW;
    a.Cout = d.Cin; a.Ho = d.H; a.Wo = d.W;
    a.KH = d.KH; a.KW = d.KW; a.pad = d.pad; a.stride = 1; a.stride_x = 1;
    int best = 5;
    long long best_cost = -1;
    for (int j = 7; j >= 1; --j) {
        const int tw = 1 << j, th = TILE_M >> j;
        const long long cost = (long long)((Wc + tw - 1) / tw) * ((Hc + th - 1) / th);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = j;
        }
    }
    a.tw_log2 = best;
    const int TW = 1 << best, TH = TILE_M >> best;
    a.tiles_x = (Wc + TW - 1) / TW;
    a.tiles_y = (Hc + TH - 1) / TH;
    a.n_cblk = (d.Cout + BLOCK_K - 1) / BLOCK_K;  // K = the forward convolution's output channels
    a.act = 0;
    a.dbg = nullptr;
    a.single_issuer = 0;
    int n_tile = 16;
    while (n_tile < d.Cin && n_tile < 128) n_tile *= 2;
    CUtensorMap mapA, mapB;
    {   // gy as (co, ox, oy, b)
        cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.y_sW * 4, (cuuint64_t)d.y_sH * 4, (cuuint64_t)d.y_sB * 4};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)TW, (cuuint32_t)TH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(gy), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the incoming gradient";
            return cudaErrorInvalidValue;
        }
    }
    {   // dgrad-packed bank as (32 k, Cin, taps * n_cblk)
        cuuint64_t dims[3] = {BLOCK_K, (cuuint64_t)d.Cin, (cuuint64_t)(d.KH * d.KW * a.n_cblk)};
        cuuint64_t strides[2] = {BLOCK_K * 4, (cuuint64_t)d.Cin * BLOCK_K * 4};
        cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)n_tile, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(w_packed), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the packed filter bank";
            return cudaErrorInvalidValue;
        }
    }
    const int n_mtiles = d.B * a.tiles_x * a.tiles_y, n_ntiles = (d.Cin + n_tile - 1) / n_tile;
    switch (n_tile) {
        case 16: return launch_dgrad_s2<16>(mapA, mapB, a, plan, n_mtiles, n_ntiles, st);
        case 32: return launch_dgrad_s2<32>(mapA, mapB, a, plan, n_mtiles, n_ntiles, st);
        case 64: return launch_dgrad_s2<64>(mapA, mapB, a, plan, n_mtiles, n_ntiles, st);
        default: return launch_dgrad_s2<128>(mapA, mapB, a, plan, n_mtiles, n_ntiles, st);
    }
}

}  // namespace tc
}  // namespace mvf
