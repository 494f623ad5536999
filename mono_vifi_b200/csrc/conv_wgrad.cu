// Weight gradient of a convolution on the tcgen05 tensor cores (TF32 inputs, fp32 accumulation in TMEM), channels-last
// activations:
//
//   dw[co,ci,kh,kw] = sum_{b,oy,ox} gy[b,oy,ox,co] * x[b, oy*s+kh-pad, ox*s+kw-pad, ci]
//
// GEMM view per CTA: for every filter tap t, D_t[128 couts, 32 cins] += A^T . B_t over the CTA's share of the output
// pixels (split-K): A = a patch of gy, [32 pixels][128 couts], B_t = the patch of x the tap sees, [32 pixels][32 cins].
// Pixels are the reduction dimension and channels are contiguous in memory, so BOTH operands are MN-major TF32
// operands (128B swizzle of 32-byte units, the only layout the tensor core accepts for them; tc_common.cuh) written
// by TMA; out-of-image pixels and channels past Cout / Cin are zero-filled by the TMA unit.  The taps' accumulators
// sit side by side in TMEM (taps x 32 columns <= 512).  Each CTA writes its partial sums to a workspace; a second
// kernel adds the splits in a fixed order (deterministic) and stores dw in the [Cout,Cin,KH,KW] layout of the
// parameter.  Replaces cuDNN's convolution_backward_weight behind nn.Conv2d of the reference's networks.
#include <cstdlib>
#include <mutex>

#include "conv_tc.cuh"
#include "pdl.cuh"
#include "tc_common.cuh"

namespace mvf {
namespace tc {

namespace {

constexpr int M_TILE = 128;  // couts per CTA (UMMA M)
constexpr int N_TILE = 32;   // cins per CTA (UMMA N), one 128-byte MN atom
constexpr int KP = 32;       // pixels per pipeline stage (4 MMAs of K = 8 per tap)
constexpr int MAX_TAPS = 9;  // taps x 32 columns <= 512 TMEM columns, 4 stages of (16 + 4 taps) KB of shared memory
constexpr int NTHREADS = 256;  // warp 0 TMA producer, warps 1 / 6 / 7 MMA issuers, warps 2..5 epilogue
constexpr int MAX_ISSUERS = 3;

struct WgradArgs {
    float* partial;  // [ksplit][taps][Cout_pad128][Cin_pad32]
    int Cout, Cin, KH, KW, pad, stride, stride_x;
    int Ho, Wo, B;
    int pw, ph;      // pixel patch of one pipeline stage: pw x ph output pixels (pw * ph a multiple of 8, <= 64)
    int patches_x, patches_y;
    int n_patches;   // B * patches_y * patches_x
    int ksplit;
    int cout_pad, cin_pad;
    int fuse_row;      // > 0: taps per MMA (N = fuse_row * 32): KW with a shared patch, a divisor of taps <= 8 otherwise
    int shared_patch;  // 1 (stride 1, ph == 1): ONE haloed x patch per stage serves all taps; 0: one x box per tap
    int pwx;           // shared patch: its pitch in pixels (pw + KW - 1)
    int stage_bytes, a_bytes, b_bytes;  // stage stride (1024-aligned), bytes landed for A and for B (all taps)
    int stages;        // pipeline depth (2..6), as many as fit in shared memory
    int max_issuers;   // MMA-issuing warps in use (1..3; MVF_WGRAD_ISSUERS for A/B)
    int g_ragged;      // 1: Cout > 32 and not a multiple of 32 (HRNet's 36 / 72 / 144): mapG is the 4-D (co, ox, oy, b) map, one box per
                       // 32-cout block, channels past Cout zero-filled by TMA
};

constexpr int MAX_STAGES = 6;

__global__ void __launch_bounds__(NTHREADS) conv_wgrad_kernel(const __grid_constant__ CUtensorMap mapG,
                                                              const __grid_constant__ CUtensorMap mapX, const WgradArgs p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int taps = p.KH * p.KW;
    const int STAGES = p.stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * p.stage_bytes);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.x * M_TILE, ci0 = blockIdx.y * N_TILE, split = blockIdx.z;
    // this CTA's contiguous share of the pixel patches
    const int per = (p.n_patches + p.ksplit - 1) / p.ksplit;
    const int k_begin = split * per, k_end = min(p.n_patches, k_begin + per);
    const int n_iters = max(0, k_end - k_begin);
    const int kp = p.pw * p.ph;  // pixels per stage
    // MMA issuers: one thread sustains one tcgen05.mma per ~70 cycles however small the MMA is (an N = 96 MMA is 48 cycles of tensor
    // work), so the tap groups of a stage -- each with its own accumulator columns -- are dealt to up to three issuing warps: group g
    // belongs to issuer g % n_issuers.  The order of additions into any accumulator is unchanged (bitwise repeatable results).
    const int n_groups = p.fuse_row ? taps / p.fuse_row : 1;
    const int n_issuers = p.fuse_row ? (n_groups < p.max_issuers ? n_groups : p.max_issuers) : 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapG);
        tma_prefetch_desc(&mapX);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], n_issuers);   // every issuer commits once per stage
        }
        mbar_init(tmem_full_bar, n_issuers);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
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
            int px = k_begin % p.patches_x, py = (k_begin / p.patches_x) % p.patches_y, b = k_begin / (p.patches_x * p.patches_y);
            for (int it = 0; it < n_iters; ++it) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int ox0 = px * p.pw, oy0 = py * p.ph;
                unsigned char* st = smem + s * p.stage_bytes;
                mbar_arrive_expect_tx(&full_bar[s], p.a_bytes + p.b_bytes);
                // A: gy as (co%32, ox, oy, b, co/32), box (32, pw, ph, 1, 4) -> smem [co/32][pixel][32 co]
                if (!p.g_ragged) {
                    tma_load_5d(st, &mapG, &full_bar[s], 0, ox0, oy0, b, co0 / 32);
                } else {
                    for (int blk = 0; blk < M_TILE / 32; ++blk)   // same shared-memory layout; a block past Cout lands as zeros
                        tma_load_4d(st + blk * (kp * 128), &mapG, &full_bar[s], co0 + 32 * blk, ox0, oy0, b);
                }
                if (p.shared_patch) {
                    // x as (ci, ix, iy, b), box (32, pw + KW - 1, KH, 1): the haloed patch all taps read
                    tma_load_4d(st + p.a_bytes, &mapX, &full_bar[s], ci0, ox0 - p.pad, oy0 - p.pad, b);
                } else {
                    // one box per tap, walked with the convolution stride -> smem [tap][pixel][32 ci]
                    int kh = 0, kw = 0;
                    for (int t = 0; t < taps; ++t) {
                        tma_load_4d(st + p.a_bytes + t * (kp * 128), &mapX, &full_bar[s], ci0, ox0 * p.stride_x + kw - p.pad,
                                    oy0 * p.stride + kh - p.pad, b);
                        if (++kw == p.KW) { kw = 0; ++kh; }
                    }
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
                if (++px == p.patches_x) { px = 0; if (++py == p.patches_y) { py = 0; ++b; } }
            }
        }
    } else if (warp == 1 || warp >= 6) {
        const int issuer = warp == 1 ? 0 : warp - 5;   // 0, 1, 2
        if (issuer < n_issuers) {
        // the whole warp runs the (uniform) loop; one elected lane issues
        constexpr uint32_t idesc = make_idesc_tf32(M_TILE, N_TILE, /*A MN-major*/ 1, /*B MN-major*/ 1);
        // MN-major, 32-byte-unit 128B swizzle: 512-byte atoms of (4 pixels x 32 channels); the next 4 pixels +512 B (SBO),
        // the next 32 couts +kp*128 B (LBO, A only: B is one atom wide)
        const uint64_t a_hi = make_smem_desc(0, kp * 128, 512, SWZ_128B_BASE32B);
        const uint64_t b_hi = make_smem_desc(0, 512, 512, SWZ_128B_BASE32B);
        const uint32_t smem_u = smem_u32(smem);
        const int kmma = kp / 8;
        int s = 0, ph = 0;
        for (int it = 0; it < n_iters; ++it) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            // descriptor = {hi word (constant), lo word}: the lo word holds the start address (14 bits) below the LBO field,
            // so stepping through the stage is a plain 32-bit add on it
            const uint32_t a_lo = ((smem_u + (uint32_t)(s * p.stage_bytes)) >> 4) | (uint32_t)a_hi;
            const uint32_t b_lo = (((smem_u + (uint32_t)(s * p.stage_bytes)) >> 4) + ((uint32_t)p.a_bytes >> 4)) | (uint32_t)b_hi;
            const uint64_t a_up = a_hi & 0xFFFFFFFF00000000ull, b_up = b_hi & 0xFFFFFFFF00000000ull;
            if (p.fuse_row) {
                // One MMA per GROUP of `fuse_row` taps.  Shared patch: the KW taps of a filter row read the x patch
                // shifted by one pixel (128 bytes) each; per-tap boxes: consecutive taps are kp * 128 bytes apart.  Either
                // way the group is `fuse_row` 32-channel atoms of ONE MN-major B operand with a constant atom stride
                // (LBO), N = fuse_row * 32.  The tensor core accepts one tcgen05.mma per ~70 cycles however small it is:
                // an N = 32 MMA is 16 cycles of work (ncu: tensor pipe 21 % active), an N = 96 one 48.  Accumulator
                // columns come out in the same order (tap * 32 + ci) as with one MMA per tap.
                if (elect_one()) {
                    const uint32_t idesc_row = make_idesc_tf32(M_TILE, p.fuse_row * N_TILE, 1, 1);
                    const uint32_t lbo = p.shared_patch ? 128u : (uint32_t)kp * 128u;
                    const uint64_t brow = make_smem_desc(0, lbo, 512, SWZ_128B_BASE32B);
                    const uint64_t brow_up = brow & 0xFFFFFFFF00000000ull;
                    const uint32_t brow_lo = (b_lo & 0x3FFFu) | (uint32_t)(brow & 0xFFFFC000ull);
                    for (int kh = issuer; kh < n_groups; kh += n_issuers) {
                        const uint32_t bt = brow_lo + (uint32_t)(p.shared_patch ? kh * p.pwx * 8 : kh * p.fuse_row * kp * 8);
                        const uint32_t dcol = tmem_d + (uint32_t)(kh * p.fuse_row * N_TILE);
#pragma unroll 4
                        for (int kk = 0; kk < kmma; ++kk)
                            umma_tf32(dcol, a_up | (uint64_t)(a_lo + kk * 64), brow_up | (uint64_t)(bt + kk * 64), idesc_row,
                                      (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                }
            } else if (elect_one()) {
                int kh = 0, kw = 0;
                for (int t = 0; t < taps; ++t) {
                    // tap t: its own box, or the shared patch shifted by kh rows and kw pixels (whole 128-byte rows)
                    const uint32_t bt = p.shared_patch ? b_lo + (uint32_t)((kh * p.pwx + kw) * 8) : b_lo + (uint32_t)(t * kp * 8);
                    const uint32_t dcol = tmem_d + (uint32_t)(t * N_TILE);
#pragma unroll 4
                    for (int kk = 0; kk < kmma; ++kk)
                        umma_tf32(dcol, a_up | (uint64_t)(a_lo + kk * 64), b_up | (uint64_t)(bt + kk * 64), idesc,
                                  (it > 0 || kk > 0) ? 1u : 0u);
                    if (++kw == p.KW) { kw = 0; ++kh; }
                }
                umma_commit(&empty_bar[s]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tmem_full_bar);
        __syncwarp();
        }
    } else {
        // epilogue: thread = one cout row; per tap 32 contiguous cins
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        if (n_iters > 0) {
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
        }
        for (int t = 0; t < taps; ++t) {
            uint32_t r[32];
            if (n_iters > 0) {
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * N_TILE), r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (co >= p.Cout) continue;  // rows past Cout are never read by the reduction
            float* dst = p.partial + (((size_t)split * taps + t) * p.cout_pad + co) * p.cin_pad + ci0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                  __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, 512);
    }
}

// dw[co][ci][tap] = sum over splits of partial[split][tap][co][ci], in ONE fixed order (bitwise reproducible).
// A CTA of 256 threads owns G = 256 / SL consecutive groups of 4 cins and SL "split lanes": thread (sl, g) adds splits sl, sl + SL, ...
// of its group -- the G threads of one split lane read G * 16 contiguous bytes of one partial, so every sector is fully used and the
// loads of a thread are independent (the round-1 kernels issued one dependent chain per output, or one warp per output whose lanes read
// 32 different sectors: 42 us per launch on the narrow high-resolution layers) -- then the SL partial sums meet in shared memory and are
// added in split-lane order.  SL is picked by the host so that narrow layers still fill the GPU.
template <int SL>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Cout, int Cin,
                                                           int taps, int ksplit, int cout_pad, int cin_pad) {
    pdl_sync();
    constexpr int G = 256 / SL;
    __shared__ float4 red[256];
    const int cin4 = (Cin + 3) / 4;
    const long long total = (long long)Cout * taps * cin4;
    const size_t split_stride = (size_t)taps * cout_pad * cin_pad;
    const int g = threadIdx.x % G, sl = threadIdx.x / G;
    for (long long base = (long long)blockIdx.x * G; base < total; base += (long long)gridDim.x * G) {
        const long long i = base + g;
        const bool ok = i < total;
        int c4 = 0, t = 0, co = 0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
            // groups are ordered (tap, cout, cin/4): consecutive groups are consecutive in the partial buffers
            c4 = (int)(i % cin4);
            const long long r = i / cin4;
            co = (int)(r % Cout);
            t = (int)(r / Cout);
            const float* src = partial + ((size_t)t * cout_pad + co) * cin_pad + 4 * c4;
#pragma unroll 4
            for (int s = sl; s < ksplit; s += SL) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)s * split_stride));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        if (SL > 1) {
            red[threadIdx.x] = acc;
            __syncthreads();
            if (sl == 0) {
#pragma unroll 4
                for (int q = 1; q < SL; ++q) {
                    const float4 v = red[q * G + g];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
        }
        if (ok && sl == 0) {
            const int ci = 4 * c4;
            float* o = dw + ((long long)co * Cin + ci) * taps + t;
            o[0] = acc.x;
            if (ci + 1 < Cin) o[taps] = acc.y;
            if (ci + 2 < Cin) o[2 * taps] = acc.z;
            if (ci + 3 < Cin) o[3 * taps] = acc.w;
        }
        if (SL > 1) __syncthreads();
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wgrad_encode_fn() {
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

int out_size(int n, int k, int pad, int stride) { return (n + 2 * pad - k) / stride + 1; }

struct Plan {
    int Ho, Wo, pw, ph, patches_x, patches_y, n_patches, ksplit, cout_pad, cin_pad, stages;
    int shared_patch, pwx, a_bytes, b_bytes, stage_bytes;
};

Plan make_plan(const WgradDesc& d) {
    Plan pl;
    pl.Ho = out_size(d.H, d.KH, d.pad, d.stride);
    pl.Wo = out_size(d.W, d.KW, d.pad, d.stride_x);
    const int taps = d.KH * d.KW;
    pl.shared_patch = (d.stride == 1 && d.stride_x == 1 && taps > 1 && !getenv("MVF_WGRAD_NO_PATCH")) ? 1 : 0;
    if (pl.shared_patch) {
        // one output-row segment of pw pixels per stage (pw a multiple of 8, <= 64): least padding, then the widest
        int best = 32;
        long long best_cost = -1;
        for (int pw = 64; pw >= 8; pw -= 8) {
            const long long cost = (long long)((pl.Wo + pw - 1) / pw) * pw;
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = pw;
            }
        }
        pl.pw = best;
        pl.ph = 1;
        pl.pwx = pl.pw + d.KW - 1;
        pl.b_bytes = d.KH * pl.pwx * 128;
    } else {
        int best = 5;
        long long best_cost = -1;
        for (int j = 5; j >= 0; --j) {
            const int pw = 1 << j, ph = KP >> j;
            const long long cost = (long long)((pl.Wo + pw - 1) / pw) * ((pl.Ho + ph - 1) / ph);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = j;
            }
        }
        pl.pw = 1 << best;
        pl.ph = KP >> best;
        pl.pwx = 0;
        pl.b_bytes = taps * pl.pw * pl.ph * 128;
    }
    pl.a_bytes = pl.pw * pl.ph * 128 * 4;
    // the MMAs of the last taps of a shared patch read up to KW - 1 rows past it: keep them inside the stage
    pl.stage_bytes = (pl.a_bytes + pl.b_bytes + (pl.shared_patch ? (d.KW - 1) * 128 : 0) + 1023) / 1024 * 1024;
    pl.patches_x = (pl.Wo + pl.pw - 1) / pl.pw;
    pl.patches_y = (pl.Ho + pl.ph - 1) / pl.ph;
    pl.n_patches = d.B * pl.patches_x * pl.patches_y;
    pl.cout_pad = (d.Cout + M_TILE - 1) / M_TILE * M_TILE;
    pl.cin_pad = (d.Cin + N_TILE - 1) / N_TILE * N_TILE;
    const int tiles = (pl.cout_pad / M_TILE) * (pl.cin_pad / N_TILE);
    // split the pixels so that about one wave of CTAs exists, but keep at least 8 patches per CTA
    static int target_ctas = 0;
    if (target_ctas == 0) {
        const char* e = std::getenv("MVF_WGRAD_CTAS");
        target_ctas = e ? std::atoi(e) : 148;
        if (target_ctas < 1) target_ctas = 148;
    }
    int ks = target_ctas / tiles;  // floor: tiles * ks <= one wave (a 150-CTA grid would cost a second wave for 2 CTAs)
    const int max_ks = pl.n_patches / 8 > 0 ? pl.n_patches / 8 : 1;
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    pl.ksplit = ks;
    pl.stages = (224 * 1024) / pl.stage_bytes;
    if (pl.stages > MAX_STAGES) pl.stages = MAX_STAGES;
    if (pl.stages < 2) pl.stages = 2;
    return pl;
}

}  // namespace

const char* wgrad_check(const WgradDesc& d) {
    if (d.B <= 0 || d.Cin <= 0 || d.H <= 0 || d.W <= 0 || d.Cout <= 0 || d.KH <= 0 || d.KW <= 0) return "non-positive size";
    if ((d.stride != 1 && d.stride != 2) || (d.stride_x != 1 && d.stride_x != 2)) return "strides must be 1 or 2";
    if (d.KH * d.KW > MAX_TAPS) return "more than 9 filter taps";
    if (d.Cin % 4 != 0 || d.Cout % 4 != 0) return "channel counts must be multiples of 4 (TMA: 16-byte pixels)";
    if ((d.x_sH % 4) || (d.x_sW % 4) || (d.x_sB % 4) || (d.g_sH % 4) || (d.g_sW % 4) || (d.g_sB % 4))
        return "strides must be multiples of 4 elements (TMA: 16 bytes)";
    if (d.pad < 0 || d.H + 2 * d.pad < d.KH || d.W + 2 * d.pad < d.KW) return "bad padding";
    return nullptr;
}

size_t wgrad_workspace_floats(const WgradDesc& d) {
    const Plan pl = make_plan(d);
    return (size_t)pl.ksplit * d.KH * d.KW * pl.cout_pad * pl.cin_pad;
}

cudaError_t conv_wgrad(const WgradDesc& d, const float* x, const float* gy, float* dw, float* workspace, cudaStream_t st,
                       const char** why) {
    *why = nullptr;
    EncodeTiledFn enc = wgrad_encode_fn();
    if (!enc) {
        *why = "cuTensorMapEncodeTiled is not available from the driver";
        return cudaErrorNotSupported;
    }
    if (((uintptr_t)x & 15) || ((uintptr_t)gy & 15) || ((uintptr_t)workspace & 15)) {
        *why = "x / grad_out / workspace pointers must be 16-byte aligned";
        return cudaErrorInvalidValue;
    }
    const Plan pl = make_plan(d);
    const int PW = pl.pw, PH = pl.ph;
    WgradArgs a;
    a.partial = workspace;
    a.Cout = d.Cout; a.Cin = d.Cin; a.KH = d.KH; a.KW = d.KW; a.pad = d.pad; a.stride = d.stride; a.stride_x = d.stride_x;
    a.Ho = pl.Ho; a.Wo = pl.Wo; a.B = d.B;
    a.pw = pl.pw; a.ph = pl.ph; a.patches_x = pl.patches_x; a.patches_y = pl.patches_y; a.n_patches = pl.n_patches;
    a.stages = pl.stages;
    a.max_issuers = MAX_ISSUERS;
    if (const char* e = getenv("MVF_WGRAD_ISSUERS")) a.max_issuers = atoi(e) < 1 ? 1 : (atoi(e) > MAX_ISSUERS ? MAX_ISSUERS : atoi(e));
    a.fuse_row = 0;
    if (!getenv("MVF_WGRAD_NO_FUSE")) {
        if (pl.shared_patch) {
            if (d.KW > 1 && d.KW * N_TILE <= 256) a.fuse_row = d.KW;
        } else if (pl.pw * pl.ph * 128 <= (0x3FFF << 4)) {  // the atom stride must fit the 14-bit LBO field
            for (int g = 8; g > 1; --g)
                if ((d.KH * d.KW) % g == 0) { a.fuse_row = g; break; }
        }
    }
    a.shared_patch = pl.shared_patch; a.pwx = pl.pwx; a.stage_bytes = pl.stage_bytes; a.a_bytes = pl.a_bytes; a.b_bytes = pl.b_bytes;
    a.ksplit = pl.ksplit; a.cout_pad = pl.cout_pad; a.cin_pad = pl.cin_pad;

    CUtensorMap mapG, mapX;
    a.g_ragged = (d.Cout > 32 && d.Cout % 32 != 0) ? 1 : 0;
    if (a.g_ragged) {   // grad_out as (co, ox, oy, b): the kernel issues one box per 32-cout block
        cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)pl.Wo, (cuuint64_t)pl.Ho, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.g_sW * 4, (cuuint64_t)d.g_sH * 4, (cuuint64_t)d.g_sB * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)PW, (cuuint32_t)PH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (enc(&mapG, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(gy), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for grad_out";
            return cudaErrorInvalidValue;
        }
    } else {   // grad_out as (co%32, ox, oy, b, co/32)
        const int c32 = d.Cout < 32 ? d.Cout : 32;
        cuuint64_t dims[5] = {(cuuint64_t)c32, (cuuint64_t)pl.Wo, (cuuint64_t)pl.Ho, (cuuint64_t)d.B, (cuuint64_t)((d.Cout + 31) / 32)};
        cuuint64_t strides[4] = {(cuuint64_t)d.g_sW * 4, (cuuint64_t)d.g_sH * 4, (cuuint64_t)d.g_sB * 4, 128};
        cuuint32_t box[5] = {32, (cuuint32_t)PW, (cuuint32_t)PH, 1, 4};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        if (enc(&mapG, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(gy), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for grad_out";
            return cudaErrorInvalidValue;
        }
    }
    {   // x as (ci, ix, iy, b), walked with the convolution stride
        cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.x_sW * 4, (cuuint64_t)d.x_sH * 4, (cuuint64_t)d.x_sB * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)(PW * d.stride_x), (cuuint32_t)(PH * d.stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)d.stride_x, (cuuint32_t)d.stride, 1};
        if (pl.shared_patch) {
            box[1] = (cuuint32_t)pl.pwx;
            box[2] = (cuuint32_t)d.KH;
        }
        if (enc(&mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the input";
            return cudaErrorInvalidValue;
        }
    }
    const int taps = d.KH * d.KW;
    const int smem = pl.stages * pl.stage_bytes + 1024 + 256;
    dim3 grid(pl.cout_pad / M_TILE, pl.cin_pad / N_TILE, pl.ksplit);
    cudaError_t e;
    static bool attr_set = false;
    if (!attr_set) {
        e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    e = launch_pdl(conv_wgrad_kernel, grid, dim3(NTHREADS), (size_t)smem, st, mapG, mapX, a);
    if (e != cudaSuccess) return e;
    // split lanes: as few as keep ~2 CTAs per SM busy, but never more lanes than splits
    const long long total = (long long)d.Cout * taps * ((d.Cin + 3) / 4);
    int sl = 1;
    while (sl < 64 && sl * 2 <= pl.ksplit && total * sl / 256 < 2 * 148) sl *= 2;
    const long long groups_per_cta = 256 / sl;
    const long long nb = (total + groups_per_cta - 1) / groups_per_cta;
    const dim3 rgrid((unsigned)(nb < 148 * 8 ? nb : 148 * 8));
#define MVF_REDUCE(SLV)                                                                                                          \
    case SLV:                                                                                                                    \
        return launch_pdl(wgrad_reduce_kernel<SLV>, rgrid, dim3(256), 0, st, (const float*)workspace, dw, d.Cout, d.Cin, taps, pl.ksplit, \
                          pl.cout_pad, pl.cin_pad)
    switch (sl) {
        MVF_REDUCE(1);
        MVF_REDUCE(2);
        MVF_REDUCE(4);
        MVF_REDUCE(8);
        MVF_REDUCE(16);
        MVF_REDUCE(32);
        default:
            return launch_pdl(wgrad_reduce_kernel<64>, rgrid, dim3(256), 0, st, (const float*)workspace, dw, d.Cout, d.Cin, taps, pl.ksplit,
                              pl.cout_pad, pl.cin_pad);
    }
#undef MVF_REDUCE
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace mvf
