// Training-mode BatchNorm2d fused with the ReLU (and the residual add) that follows it in the ResNet blocks, on dense
// channels-last tensors [P = B*H*W pixels][C channels] (torchvision BasicBlock / Bottleneck as used by
// networks/monodepth2.py:16-31 and networks/posenet.py:10-52; hrnet_encoder.py:58-139):
//
//   forward : y = relu( (x - mean_c) * rsqrt(var_c + eps) * gamma_c + beta_c  [+ identity] )
//   backward: g = grad_y * (y > 0);  d_beta = sum g;  d_gamma = sum g * xhat;
//             grad_x = gamma * invstd * (g - d_beta / P - xhat * d_gamma / P);  grad_identity = g
//
// HBM-bound: the forward reads x twice (statistics, apply) [+ identity once] and writes y once; the backward reads
// (grad_y, y, x) twice and writes grad_x [+ grad_identity].  Per-channel sums are accumulated per CTA, written to a
// workspace and added in a fixed order by a small finalize kernel (double precision): bitwise deterministic.
#include "bn_cl.cuh"

#include <cstdlib>
#include "pdl.cuh"

namespace mvf {
namespace {

constexpr int NT = 256;

// partial[block][0][c], partial[block][1][c]: two per-channel sums of this block's pixel share
template <bool BWD>
__global__ void __launch_bounds__(NT) bn_partial_kernel(const float4* __restrict__ x, const float4* __restrict__ gy,
                                                        const float4* __restrict__ y, const float* __restrict__ mean,
                                                        const float* __restrict__ invstd, float* __restrict__ partial, long long P,
                                                        int C4, int relu) {
    pdl_sync();
    extern __shared__ float4 red[];  // [rows][C4] x 2
    const int rows = NT / C4 > 0 ? NT / C4 : 1;  // threads are (row, channel group); C4 <= NT is checked by the host
    const int c = threadIdx.x % C4, r = threadIdx.x / C4;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (r < rows) {
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f), is = m;
        if (BWD) {
            m = *reinterpret_cast<const float4*>(mean + 4 * c);
            is = *reinterpret_cast<const float4*>(invstd + 4 * c);
        }
        // four pixels per trip so that four (twelve in the backward) independent 16-byte loads are in flight per thread
        const long long step = (long long)gridDim.x * rows;
        for (long long p0 = (long long)blockIdx.x * rows + r; p0 < P; p0 += 4 * step) {
            float4 v[4], g[4], o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long p = p0 + u * step;
                const bool ok = p < P;
                v[u] = ok ? __ldg(x + p * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (BWD) {
                    g[u] = ok ? __ldg(gy + p * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    o[u] = (ok && relu) ? __ldg(y + p * C4 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
                    if (!ok) v[u] = m;  // xhat = 0
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!BWD) {
                    s0.x += v[u].x; s0.y += v[u].y; s0.z += v[u].z; s0.w += v[u].w;
                    s1.x = fmaf(v[u].x, v[u].x, s1.x); s1.y = fmaf(v[u].y, v[u].y, s1.y);
                    s1.z = fmaf(v[u].z, v[u].z, s1.z); s1.w = fmaf(v[u].w, v[u].w, s1.w);
                } else {
                    float4 q = g[u];
                    q.x = o[u].x > 0.f ? q.x : 0.f; q.y = o[u].y > 0.f ? q.y : 0.f; q.z = o[u].z > 0.f ? q.z : 0.f; q.w = o[u].w > 0.f ? q.w : 0.f;
                    s0.x += q.x; s0.y += q.y; s0.z += q.z; s0.w += q.w;
                    s1.x = fmaf(q.x, (v[u].x - m.x) * is.x, s1.x); s1.y = fmaf(q.y, (v[u].y - m.y) * is.y, s1.y);
                    s1.z = fmaf(q.z, (v[u].z - m.z) * is.z, s1.z); s1.w = fmaf(q.w, (v[u].w - m.w) * is.w, s1.w);
                }
            }
        }
        red[r * C4 + c] = s0;
        red[(rows + r) * C4 + c] = s1;
    }
    __syncthreads();
    if (threadIdx.x < C4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        for (int q = 0; q < rows; ++q) {  // fixed order
            const float4 u = red[q * C4 + threadIdx.x], w = red[(rows + q) * C4 + threadIdx.x];
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += w.x; b.y += w.y; b.z += w.z; b.w += w.w;
        }
        float4* out = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * 2 * C4;
        out[threadIdx.x] = a;
        out[C4 + threadIdx.x] = b;
    }
}

// Sum of the per-CTA partials of one channel, in a fixed order.  A CTA of 32 x FIN_BL threads owns 32 consecutive channels: thread
// (cl, bl) adds blocks bl, bl + FIN_BL, ... of channel c0 + cl -- for a given block the 32 channels are one coalesced 128-byte read
// (round 1 had one warp per channel whose lanes read 32 different cache lines per load: 6-15 us per launch, on the critical path of
// every BatchNorm) -- and the FIN_BL partial sums meet in shared memory.  Returns true in the thread that holds the result.
// The kernel is a pure latency chain (a 64-channel layer launches TWO CTAs), so the walk over the blocks is as short and as wide as it
// gets: 32 block lanes (1024 threads) and four blocks per trip with their loads issued together.  With 8 lanes and one block per trip a
// thread made 45 dependent round trips to L2 for 360 blocks: 32 us per launch under ncu (17 us in the step, profiles/r2_launches_summary.txt),
// 120 launches per step, each between a BatchNorm's statistics and its apply kernel.
constexpr int FIN_BL = 32;
constexpr int FIN_UNROLL = 4;
__device__ __forceinline__ bool channel_sums(const float* __restrict__ partial, int nblocks, int C, int fold, int& c, double& s, double& ss) {
    __shared__ double red[FIN_BL][2][32];
    const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
    c = blockIdx.x * 32 + cl;
    const int Cv = fold * C;
    s = 0.0;
    ss = 0.0;
    if (c < C)
        for (int b0 = bl; b0 < nblocks; b0 += FIN_UNROLL * FIN_BL) {
            float v[FIN_UNROLL][2][2];
#pragma unroll
            for (int u = 0; u < FIN_UNROLL; ++u) {
                const int b = b0 + u * FIN_BL;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const bool ok = b < nblocks && j < fold;
                    v[u][j][0] = ok ? partial[(size_t)b * 2 * Cv + j * C + c] : 0.f;
                    v[u][j][1] = ok ? partial[(size_t)b * 2 * Cv + Cv + j * C + c] : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < FIN_UNROLL; ++u)
#pragma unroll
                for (int j = 0; j < 2; ++j) {   // (fold <= 2; the slots past it hold zeros)
                    s += (double)v[u][j][0];
                    ss += (double)v[u][j][1];
                }
        }
    red[bl][0][cl] = s;
    red[bl][1][cl] = ss;
    __syncthreads();
    if (bl != 0 || c >= C) return false;
#pragma unroll 4
    for (int q = 1; q < FIN_BL; ++q) {
        s += red[q][0][cl];
        ss += red[q][1][cl];
    }
    return true;
}

// forward: mean / biased variance -> (mean, invstd) saved for backward, running statistics updated as nn.BatchNorm2d does.
// fold > 1 (channel counts with C % 4 == 2, HRNet's 18): the tensor [P][C] is processed as [P / fold][fold * C] so that the float4
// kernels apply; "virtual" channel j * C + c is real channel c at pixels of parity j.  The partial sums of the fold virtual channels of a
// real channel are added here, mean / invstd (and copies of gamma / beta) are written for every virtual channel, the running
// statistics once per real channel.  C is the REAL channel count, P the real pixel count.
__global__ void __launch_bounds__(32 * FIN_BL) bn_finalize_fwd_kernel(const float* __restrict__ partial, int nblocks, int C, long long P, float eps, float momentum,
                                       float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, long long* __restrict__ num_batches_tracked, int fold,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ gamma_v,
                                       float* __restrict__ beta_v) {
    pdl_sync();
    if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
    int c;
    double s, ss;
    if (!channel_sums(partial, nblocks, C, fold, c, s, ss)) return;
    const double m = s / (double)P;
    double var = ss / (double)P - m * m;
    if (var < 0.0) var = 0.0;
    for (int j = 0; j < fold; ++j) {
        mean[j * C + c] = (float)m;
        invstd[j * C + c] = (float)(1.0 / sqrt(var + (double)eps));
        if (gamma_v) {
            gamma_v[j * C + c] = gamma[c];
            beta_v[j * C + c] = beta[c];
        }
    }
    if (running_mean) {
        const double unbiased = P > 1 ? var * (double)P / (double)(P - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// backward: d_beta = sum g, d_gamma = sum g * xhat (fold: summed over the virtual channels of a real channel; per-virtual-channel
// copies of d_gamma / d_beta / gamma for the apply kernel)
__global__ void __launch_bounds__(32 * FIN_BL) bn_finalize_bwd_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int fold, const float* __restrict__ gamma, float* __restrict__ dgamma_v,
                                       float* __restrict__ dbeta_v, float* __restrict__ gamma_v) {
    pdl_sync();
    int c;
    double s, ss;
    if (!channel_sums(partial, nblocks, C, fold, c, s, ss)) return;
    dbeta[c] = (float)s;
    dgamma[c] = (float)ss;
    if (dgamma_v)
        for (int j = 0; j < fold; ++j) {
            dgamma_v[j * C + c] = (float)ss;
            dbeta_v[j * C + c] = (float)s;
            gamma_v[j * C + c] = gamma[c];
        }
}

__global__ void bn_apply_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ identity, float4* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, long long total4, int C4, int relu) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = total4 < 0xffffffffLL ? (int)((unsigned)i % (unsigned)C4) : (int)(i % C4);
        const float4 m = *reinterpret_cast<const float4*>(mean + 4 * c), is = *reinterpret_cast<const float4*>(invstd + 4 * c);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + 4 * c), be = *reinterpret_cast<const float4*>(beta + 4 * c);
        const float4 v = __ldg(x + i);
        float4 o;
        o.x = fmaf((v.x - m.x) * is.x, ga.x, be.x); o.y = fmaf((v.y - m.y) * is.y, ga.y, be.y);
        o.z = fmaf((v.z - m.z) * is.z, ga.z, be.z); o.w = fmaf((v.w - m.w) * is.w, ga.w, be.w);
        if (identity) {
            const float4 d = __ldg(identity + i);
            o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        y[i] = o;
    }
}

__global__ void bn_apply_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, const float4* __restrict__ y,
                                    float4* __restrict__ gx, float4* __restrict__ gid, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, long long total4, int C4, float inv_P, int relu) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = total4 < 0xffffffffLL ? (int)((unsigned)i % (unsigned)C4) : (int)(i % C4);
        const float4 m = *reinterpret_cast<const float4*>(mean + 4 * c), is = *reinterpret_cast<const float4*>(invstd + 4 * c);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 dg = *reinterpret_cast<const float4*>(dgamma + 4 * c), db = *reinterpret_cast<const float4*>(dbeta + 4 * c);
        const float4 v = __ldg(x + i);
        float4 g = __ldg(gy + i);
        if (relu) {
            const float4 o = __ldg(y + i);
            g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
        }
        if (gid) gid[i] = g;
        float4 r;
        r.x = ga.x * is.x * (g.x - db.x * inv_P - (v.x - m.x) * is.x * dg.x * inv_P);
        r.y = ga.y * is.y * (g.y - db.y * inv_P - (v.y - m.y) * is.y * dg.y * inv_P);
        r.z = ga.z * is.z * (g.z - db.z * inv_P - (v.z - m.z) * is.z * dg.z * inv_P);
        r.w = ga.w * is.w * (g.w - db.w * inv_P - (v.w - m.w) * is.w * dg.w * inv_P);
        gx[i] = r;
    }
}

// ---- activation backward + bias gradient of a convolution output, one pass --------------------------------------
// gpre = gy * act'(y) (act: 0 none, 1 relu, 2 elu, both from the saved OUTPUT) and partial[block][c] = this block's
// share of sum_pixels gpre[.., c]; the finalize kernel adds the blocks in a fixed order.
__global__ void __launch_bounds__(NT) act_bias_partial_kernel(const float4* __restrict__ gy, const float4* __restrict__ y,
                                                              float4* __restrict__ gpre, float* __restrict__ partial, long long P,
                                                              int C4, int act) {
    pdl_sync();
    extern __shared__ float4 red[];  // [rows][C4]
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    const int c = threadIdx.x % C4, r = threadIdx.x / C4;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
        const long long step = (long long)gridDim.x * rows;
        for (long long p0 = (long long)blockIdx.x * rows + r; p0 < P; p0 += 4 * step) {
            float4 g[4], o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long p = p0 + u * step;
                const bool ok = p < P;
                g[u] = ok ? __ldg(gy + p * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                o[u] = (ok && act) ? __ldg(y + p * C4 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 q = g[u];
                if (act == 1) {
                    q.x = o[u].x > 0.f ? q.x : 0.f; q.y = o[u].y > 0.f ? q.y : 0.f; q.z = o[u].z > 0.f ? q.z : 0.f; q.w = o[u].w > 0.f ? q.w : 0.f;
                } else if (act == 2) {
                    q.x = o[u].x > 0.f ? q.x : q.x * (o[u].x + 1.f); q.y = o[u].y > 0.f ? q.y : q.y * (o[u].y + 1.f);
                    q.z = o[u].z > 0.f ? q.z : q.z * (o[u].z + 1.f); q.w = o[u].w > 0.f ? q.w : q.w * (o[u].w + 1.f);
                }
                const long long p = p0 + u * step;
                if (gpre && p < P) gpre[p * C4 + c] = q;
                s0.x += q.x; s0.y += q.y; s0.z += q.z; s0.w += q.w;
            }
        }
        red[r * C4 + c] = s0;
    }
    __syncthreads();
    if (partial && threadIdx.x < C4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < rows; ++q) {  // fixed order
            const float4 u = red[q * C4 + threadIdx.x];
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
        }
        reinterpret_cast<float4*>(partial)[(size_t)blockIdx.x * C4 + threadIdx.x] = a;
    }
}

__global__ void bias_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ gbias) {
    pdl_sync();
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += (double)partial[(size_t)b * C + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) gbias[c] = (float)s;
}

// ---- cross-rank (SyncBatchNorm) variants ------------------------------------------------------------------------------
// Under data parallelism the reference converts every BatchNorm to nn.SyncBatchNorm (train.py:205-208): statistics are
// taken over the batches of ALL ranks.  The per-CTA partials are first reduced to one [2C + 1] vector of doubles per
// rank {sum_0[C], sum_1[C], pixel count}; the ranks exchange and add those vectors (NCCL all-reduce, or the peer-memory
// exchange kernel below) and every rank finishes with the global sums.
//   forward : sum_0 = sum x, sum_1 = sum x^2      -> mean, biased variance over the global count
//   backward: sum_0 = sum g, sum_1 = sum g * xhat -> the two means of the input gradient over the global count; the
//             parameter gradients d_beta / d_gamma stay LOCAL sums (they are averaged with all other gradients later).
__global__ void __launch_bounds__(32 * FIN_BL) bn_sums_kernel(const float* __restrict__ partial, int nblocks, int C, double count, double* __restrict__ sums,
                               float* __restrict__ local0, float* __restrict__ local1) {
    pdl_sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) sums[2 * C] = count;
    int c;
    double s, ss;
    if (!channel_sums(partial, nblocks, C, 1, c, s, ss)) return;
    sums[c] = s;
    sums[C + c] = ss;
    if (local0) local0[c] = (float)s;     // backward: d_beta
    if (local1) local1[c] = (float)ss;    //           d_gamma
}

__global__ void bn_finalize_sync_fwd_kernel(const double* __restrict__ sums, int C, float eps, float momentum, float* __restrict__ mean,
                                            float* __restrict__ invstd, float* __restrict__ running_mean,
                                            float* __restrict__ running_var, long long* __restrict__ num_batches_tracked) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (num_batches_tracked && c == 0) *num_batches_tracked += 1;
    if (c >= C) return;
    const double N = sums[2 * C];
    const double m = sums[c] / N;
    double var = sums[C + c] / N - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unbiased = N > 1.0 ? var * N / (N - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// global sums -> what bn_apply_bwd_kernel reads: d_beta, d_gamma over all ranks and 1 / (global pixel count)
__global__ void bn_unpack_sync_bwd_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma_g, float* __restrict__ dbeta_g,
                                          float* __restrict__ inv_count) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) *inv_count = (float)(1.0 / sums[2 * C]);
    if (c >= C) return;
    dbeta_g[c] = (float)sums[c];
    dgamma_g[c] = (float)sums[C + c];
}

__global__ void bn_apply_bwd_sync_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, const float4* __restrict__ y,
                                         float4* __restrict__ gx, float4* __restrict__ gid, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                         const float* __restrict__ dbeta, long long total4, int C4, const float* __restrict__ inv_count,
                                         int relu) {
    pdl_sync();
    const float inv_P = *inv_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = total4 < 0xffffffffLL ? (int)((unsigned)i % (unsigned)C4) : (int)(i % C4);
        const float4 m = *reinterpret_cast<const float4*>(mean + 4 * c), is = *reinterpret_cast<const float4*>(invstd + 4 * c);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 dg = *reinterpret_cast<const float4*>(dgamma + 4 * c), db = *reinterpret_cast<const float4*>(dbeta + 4 * c);
        const float4 v = __ldg(x + i);
        float4 g = __ldg(gy + i);
        if (relu) {
            const float4 o = __ldg(y + i);
            g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
        }
        if (gid) gid[i] = g;
        float4 r;
        r.x = ga.x * is.x * (g.x - db.x * inv_P - (v.x - m.x) * is.x * dg.x * inv_P);
        r.y = ga.y * is.y * (g.y - db.y * inv_P - (v.y - m.y) * is.y * dg.y * inv_P);
        r.z = ga.z * is.z * (g.z - db.z * inv_P - (v.z - m.z) * is.z * dg.z * inv_P);
        r.w = ga.w * is.w * (g.w - db.w * inv_P - (v.w - m.w) * is.w * dg.w * inv_P);
        gx[i] = r;
    }
}

int partial_blocks(long long P, int C4) {
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    static int per_sm = 0, min_px = 0;
    if (per_sm == 0) {
        const char* e = std::getenv("MVF_BN_BLOCKS_PER_SM");
        per_sm = e ? std::atoi(e) : 4;
        const char* m = std::getenv("MVF_BN_MIN_PX");
        min_px = m ? std::atoi(m) : 16;
    }
    long long nb = (P + (long long)rows * min_px - 1) / ((long long)rows * min_px);  // at least ~min_px pixels per thread
    if (nb > 148 * per_sm) nb = 148 * per_sm;
    if (nb < 1) nb = 1;
    return (int)nb;
}
int apply_blocks(long long total4) {
    long long nb = (total4 + NT - 1) / NT;
    return (int)(nb > 148 * 16 ? 148 * 16 : (nb < 1 ? 1 : nb));
}

}  // namespace

// C % 4 == 2 (and an even pixel count): two pixels are one row of 2C "virtual" channels
static inline int fold_of(long long P, int C) { return (C % 4 == 0) ? 1 : 2; }

size_t bn_workspace_floats(long long P, int C) {
    const int fold = fold_of(P, C), Cv = fold * C;
    return (size_t)partial_blocks(P / fold, Cv / 4) * 2 * Cv + (fold > 1 ? 4 * (size_t)Cv : 0);
}

cudaError_t bn_forward(const float* x, const float* identity, float* y, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, long long* num_batches_tracked, float* save_mean, float* save_invstd, float* workspace,
                       long long P, int C, float eps, float momentum, int relu, cudaStream_t st) {
    // save_mean / save_invstd hold fold * C entries (the per-virtual-channel copies the backward kernels read)
    const int fold = fold_of(P, C), Cv = fold * C, C4 = Cv / 4;
    const long long Pv = P / fold;
    const int nb = partial_blocks(Pv, C4);
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    float* gamma_v = fold > 1 ? workspace + (size_t)nb * 2 * Cv : nullptr;
    float* beta_v = fold > 1 ? gamma_v + Cv : nullptr;
    launch_pdl(bn_partial_kernel<false>, dim3((unsigned)(nb)), dim3(NT), (size_t)(2 * rows * C4 * sizeof(float4)), st, (const float4*)x, nullptr, nullptr, nullptr, nullptr,
                                                                            workspace, Pv, C4, 0);
    launch_pdl(bn_finalize_fwd_kernel, dim3((unsigned)((C + 31) / 32)), dim3(32 * FIN_BL), (size_t)(0), st, (const float*)workspace, nb, C, P, eps, momentum, save_mean, save_invstd,
               running_mean, running_var, num_batches_tracked, fold, gamma, beta, gamma_v, beta_v);
    const long long total4 = Pv * C4;
    launch_pdl(bn_apply_fwd_kernel, dim3((unsigned)(apply_blocks(total4))), dim3(NT), (size_t)(0), st, (const float4*)x, (const float4*)identity, (float4*)y,
               (const float*)save_mean, (const float*)save_invstd, fold > 1 ? (const float*)gamma_v : gamma, fold > 1 ? (const float*)beta_v : beta, total4, C4, relu);
    return cudaGetLastError();
}

cudaError_t bn_backward(const float* x, const float* gy, const float* y, const float* gamma, const float* save_mean,
                        const float* save_invstd, float* gx, float* gidentity, float* dgamma, float* dbeta, float* workspace,
                        long long P, int C, int relu, cudaStream_t st) {
    const int fold = fold_of(P, C), Cv = fold * C, C4 = Cv / 4;
    const long long Pv = P / fold;
    const int nb = partial_blocks(Pv, C4);
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    float* dgamma_v = fold > 1 ? workspace + (size_t)nb * 2 * Cv : nullptr;
    float* dbeta_v = fold > 1 ? dgamma_v + Cv : nullptr;
    float* gamma_v = fold > 1 ? dbeta_v + Cv : nullptr;
    launch_pdl(bn_partial_kernel<true>, dim3((unsigned)(nb)), dim3(NT), (size_t)(2 * rows * C4 * sizeof(float4)), st, (const float4*)x, (const float4*)gy, (const float4*)y,
                                                                           save_mean, save_invstd, workspace, Pv, C4, relu);
    launch_pdl(bn_finalize_bwd_kernel, dim3((unsigned)((C + 31) / 32)), dim3(32 * FIN_BL), (size_t)(0), st, (const float*)workspace, nb, C, dgamma, dbeta, fold, gamma, dgamma_v,
               dbeta_v, gamma_v);
    const long long total4 = Pv * C4;
    launch_pdl(bn_apply_bwd_kernel, dim3((unsigned)(apply_blocks(total4))), dim3(NT), (size_t)(0), st, (const float4*)x, (const float4*)gy, (const float4*)y, (float4*)gx,
                                                           (float4*)gidentity, save_mean, save_invstd, fold > 1 ? (const float*)gamma_v : gamma,
                                                           fold > 1 ? (const float*)dgamma_v : (const float*)dgamma, fold > 1 ? (const float*)dbeta_v : (const float*)dbeta, total4,
                                                           C4, (float)(1.0 / (double)P), relu);
    return cudaGetLastError();
}

cudaError_t bn_sync_stats_fwd(const float* x, double* sums, float* workspace, long long P, int C, cudaStream_t st) {
    const int C4 = C / 4, nb = partial_blocks(P, C4);
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    launch_pdl(bn_partial_kernel<false>, dim3((unsigned)nb), dim3(NT), (size_t)(2 * rows * C4 * sizeof(float4)), st, (const float4*)x,
               nullptr, nullptr, nullptr, nullptr, workspace, P, C4, 0);
    launch_pdl(bn_sums_kernel, dim3((unsigned)((C + 31) / 32)), dim3(32 * FIN_BL), (size_t)0, st, (const float*)workspace, nb, C, (double)P, sums,
               nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t bn_sync_apply_fwd(const float* x, const float* identity, float* y, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, long long* num_batches_tracked, float* save_mean,
                              float* save_invstd, const double* sums, long long P, int C, float eps, float momentum, int relu,
                              cudaStream_t st) {
    const int C4 = C / 4;
    launch_pdl(bn_finalize_sync_fwd_kernel, dim3((unsigned)((C + 127) / 128)), dim3(128), (size_t)0, st, sums, C, eps, momentum, save_mean,
               save_invstd, running_mean, running_var, num_batches_tracked);
    const long long total4 = P * C4;
    launch_pdl(bn_apply_fwd_kernel, dim3((unsigned)(apply_blocks(total4))), dim3(NT), (size_t)0, st, (const float4*)x,
               (const float4*)identity, (float4*)y, (const float*)save_mean, (const float*)save_invstd, gamma, beta, total4, C4, relu);
    return cudaGetLastError();
}

cudaError_t bn_sync_stats_bwd(const float* x, const float* gy, const float* y, const float* save_mean, const float* save_invstd,
                              double* sums, float* dgamma, float* dbeta, float* workspace, long long P, int C, int relu,
                              cudaStream_t st) {
    const int C4 = C / 4, nb = partial_blocks(P, C4);
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    launch_pdl(bn_partial_kernel<true>, dim3((unsigned)nb), dim3(NT), (size_t)(2 * rows * C4 * sizeof(float4)), st, (const float4*)x,
               (const float4*)gy, (const float4*)y, save_mean, save_invstd, workspace, P, C4, relu);
    launch_pdl(bn_sums_kernel, dim3((unsigned)((C + 31) / 32)), dim3(32 * FIN_BL), (size_t)0, st, (const float*)workspace, nb, C, (double)P, sums,
               dbeta, dgamma);
    return cudaGetLastError();
}

// scratch: 2C + 4 floats (global d_gamma, d_beta, 1 / count)
cudaError_t bn_sync_apply_bwd(const float* x, const float* gy, const float* y, const float* gamma, const float* save_mean,
                              const float* save_invstd, float* gx, float* gidentity, const double* sums, float* scratch,
                              long long P, int C, int relu, cudaStream_t st) {
    const int C4 = C / 4;
    float* dg = scratch;
    float* db = scratch + C;
    float* inv = scratch + 2 * C;
    launch_pdl(bn_unpack_sync_bwd_kernel, dim3((unsigned)((C + 127) / 128)), dim3(128), (size_t)0, st, sums, C, dg, db, inv);
    const long long total4 = P * C4;
    launch_pdl(bn_apply_bwd_sync_kernel, dim3((unsigned)(apply_blocks(total4))), dim3(NT), (size_t)0, st, (const float4*)x,
               (const float4*)gy, (const float4*)y, (float4*)gx, (float4*)gidentity, save_mean, save_invstd, gamma, (const float*)dg,
               (const float*)db, total4, C4, (const float*)inv, relu);
    return cudaGetLastError();
}

cudaError_t act_bwd_bias(const float* gy, const float* y, float* gpre, float* gbias, float* workspace, long long P, int C, int act,
                         cudaStream_t st) {
    const int C4 = C / 4, nb = partial_blocks(P, C4);
    const int rows = NT / C4 > 0 ? NT / C4 : 1;
    launch_pdl(act_bias_partial_kernel, dim3((unsigned)nb), dim3(NT), (size_t)(rows * C4 * sizeof(float4)), st, (const float4*)gy,
               (const float4*)y, (float4*)gpre, gbias ? workspace : nullptr, P, C4, act);
    if (gbias) launch_pdl(bias_finalize_kernel, dim3((unsigned)((C + 3) / 4)), dim3(128), (size_t)0, st, (const float*)workspace, nb, C, gbias);
    return cudaGetLastError();
}

}  // namespace mvf
