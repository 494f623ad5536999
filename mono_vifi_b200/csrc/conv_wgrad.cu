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
#include <mutex>

#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace mvf {
namespace tc {

namespace {

constexpr int M_TILE = 128;  // couts per CTA (UMMA M)
constexpr int N_TILE = 32;   // cins per CTA (UMMA N), one 128-byte MN atom
constexpr int KP = 32;       // pixels per pipeline stage (4 MMAs of K = 8 per tap)
constexpr int MAX_TAPS = 9;  // taps x 32 columns <= 512 TMEM columns, 4 stages of (16 + 4 taps) KB of shared memory
constexpr int A_BYTES = M_TILE * KP * 4;      // 16 KB
constexpr int B_TAP_BYTES = N_TILE * KP * 4;  // 4 KB
constexpr int NTHREADS = 192;

struct WgradArgs {
    float* partial;  // [ksplit][taps][Cout_pad128][Cin_pad32]
    int Cout, Cin, KH, KW, pad, stride;
    int Ho, Wo, B;
    int pw_log2;     // pixel patch = (1 << pw_log2) x (32 >> pw_log2) output pixels
    int patches_x, patches_y;
    int n_patches;   // B * patches_y * patches_x
    int ksplit;
    int cout_pad, cin_pad;
};

template <int STAGES>
__global__ void __launch_bounds__(NTHREADS) conv_wgrad_kernel(const __grid_constant__ CUtensorMap mapG,
                                                              const __grid_constant__ CUtensorMap mapX, const WgradArgs p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int taps = p.KH * p.KW;
    const int stage_bytes = A_BYTES + taps * B_TAP_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.x * M_TILE, ci0 = blockIdx.y * N_TILE, split = blockIdx.z;
    // this CTA's contiguous share of the pixel patches
    const int per = (p.n_patches + p.ksplit - 1) / p.ksplit;
    const int k_begin = split * per, k_end = min(p.n_patches, k_begin + per);
    const int n_iters = max(0, k_end - k_begin);
    const int PW = 1 << p.pw_log2, PH = KP >> p.pw_log2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapG);
        tma_prefetch_desc(&mapX);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
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

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int k = k_begin + it;
                const int px = k % p.patches_x, py = (k / p.patches_x) % p.patches_y, b = k / (p.patches_x * p.patches_y);
                const int ox0 = px * PW, oy0 = py * PH;
                unsigned char* st = smem + s * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                // A: gy as (co%32, ox, oy, b, co/32), box (32, PW, PH, 1, 4) -> smem [co/32][pixel][32 co]
                tma_load_5d(st, &mapG, &full_bar[s], 0, ox0, oy0, b, co0 / 32);
                // B_t: x as (ci, ix, iy, b), box (32, PW*s, PH*s, 1) walked with the conv stride -> smem [pixel][32 ci]
                for (int t = 0; t < taps; ++t) {
                    const int kh = t / p.KW, kw = t - kh * p.KW;
                    tma_load_4d(st + A_BYTES + t * B_TAP_BYTES, &mapX, &full_bar[s], ci0, ox0 * p.stride + kw - p.pad,
                                oy0 * p.stride + kh - p.pad, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(M_TILE, N_TILE, /*A MN-major*/ 1, /*B MN-major*/ 1);
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + s * stage_bytes), b_base = a_base + A_BYTES;
                for (int t = 0; t < taps; ++t) {
#pragma unroll
                    for (int kk = 0; kk < KP / 8; ++kk) {
                        // MN-major, 32-byte-unit 128B swizzle: 512-byte atoms of (4 pixels x 32 channels); the next 4 pixels
                        // +512 B (SBO), the next 32 couts +KP*128 B (LBO); this MMA's 8 pixels start kk*1024 B in
                        const uint64_t adesc = make_smem_desc(a_base + kk * 1024, KP * 128, 512, SWZ_128B_BASE32B);
                        const uint64_t bdesc = make_smem_desc(b_base + t * B_TAP_BYTES + kk * 1024, KP * 128, 512, SWZ_128B_BASE32B);
                        umma_tf32(tmem_d + (uint32_t)(t * N_TILE), adesc, bdesc, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(tmem_full_bar);
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

// dw[co][ci][tap] = sum over splits (fixed order) of partial[split][tap][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Cout, int Cin, int taps,
                                    int ksplit, int cout_pad, int cin_pad) {
    const long long total = (long long)Cout * Cin * taps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // consecutive threads walk ci fastest so that the partial reads coalesce
        const int ci = (int)(i % Cin);
        long long r = i / Cin;
        const int t = (int)(r % taps);
        const int co = (int)(r / taps);
        float acc = 0.f;
        for (int s = 0; s < ksplit; ++s) acc += partial[(((size_t)s * taps + t) * cout_pad + co) * cin_pad + ci];
        dw[((long long)co * Cin + ci) * taps + t] = acc;
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
    int Ho, Wo, pw_log2, patches_x, patches_y, n_patches, ksplit, cout_pad, cin_pad, stages;
};

Plan make_plan(const WgradDesc& d) {
    Plan pl;
    pl.Ho = out_size(d.H, d.KH, d.pad, d.stride);
    pl.Wo = out_size(d.W, d.KW, d.pad, d.stride);
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
    pl.pw_log2 = best;
    pl.patches_x = (pl.Wo + (1 << best) - 1) >> best;
    const int ph = KP >> best;
    pl.patches_y = (pl.Ho + ph - 1) / ph;
    pl.n_patches = d.B * pl.patches_x * pl.patches_y;
    pl.cout_pad = (d.Cout + M_TILE - 1) / M_TILE * M_TILE;
    pl.cin_pad = (d.Cin + N_TILE - 1) / N_TILE * N_TILE;
    const int tiles = (pl.cout_pad / M_TILE) * (pl.cin_pad / N_TILE);
    // split the pixels so that about one wave of CTAs exists, but keep at least 8 patches per CTA
    int ks = (148 + tiles - 1) / tiles;
    const int max_ks = pl.n_patches / 8 > 0 ? pl.n_patches / 8 : 1;
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    pl.ksplit = ks;
    pl.stages = 4;
    return pl;
}

}  // namespace

const char* wgrad_check(const WgradDesc& d) {
    if (d.B <= 0 || d.Cin <= 0 || d.H <= 0 || d.W <= 0 || d.Cout <= 0 || d.KH <= 0 || d.KW <= 0) return "non-positive size";
    if (d.stride != 1 && d.stride != 2) return "stride must be 1 or 2";
    if (d.KH * d.KW > MAX_TAPS) return "more than 9 filter taps";
    if (d.Cin % 4 != 0 || d.Cout % 4 != 0) return "channel counts must be multiples of 4 (TMA: 16-byte pixels)";
    if (d.Cout > 32 && d.Cout % 32 != 0) return "Cout above 32 must be a multiple of 32";
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
    const int PW = 1 << pl.pw_log2, PH = KP >> pl.pw_log2;
    WgradArgs a;
    a.partial = workspace;
    a.Cout = d.Cout; a.Cin = d.Cin; a.KH = d.KH; a.KW = d.KW; a.pad = d.pad; a.stride = d.stride;
    a.Ho = pl.Ho; a.Wo = pl.Wo; a.B = d.B;
    a.pw_log2 = pl.pw_log2; a.patches_x = pl.patches_x; a.patches_y = pl.patches_y; a.n_patches = pl.n_patches;
    a.ksplit = pl.ksplit; a.cout_pad = pl.cout_pad; a.cin_pad = pl.cin_pad;

    CUtensorMap mapG, mapX;
    {   // grad_out as (co%32, ox, oy, b, co/32)
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
        cuuint32_t box[4] = {32, (cuuint32_t)(PW * d.stride), (cuuint32_t)(PH * d.stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
        if (enc(&mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            *why = "cuTensorMapEncodeTiled failed for the input";
            return cudaErrorInvalidValue;
        }
    }
    const int taps = d.KH * d.KW;
    const int stage_bytes = A_BYTES + taps * B_TAP_BYTES;
    const int smem = pl.stages * stage_bytes + 1024 + 256;
    dim3 grid(pl.cout_pad / M_TILE, pl.cin_pad / N_TILE, pl.ksplit);
    cudaError_t e;
    static bool attr_set = false;
    if (!attr_set) {
        e = cudaFuncSetAttribute(conv_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 4 * (A_BYTES + MAX_TAPS * B_TAP_BYTES) + 1280);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    conv_wgrad_kernel<4><<<grid, NTHREADS, smem, st>>>(mapG, mapX, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const long long total = (long long)d.Cout * d.Cin * taps;
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads < 1184 ? (total + threads - 1) / threads : 1184);
    wgrad_reduce_kernel<<<blocks, threads, 0, st>>>(workspace, dw, d.Cout, d.Cin, taps, pl.ksplit, pl.cout_pad, pl.cin_pad);
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace mvf
