// F1 forward, persistent warp-specialised variant (round 2): the same arithmetic as f1_fwd.cu --
//   disp_to_depth -> BackprojectDepth -> Project3D -> grid_sample(border, align_corners) of 2 sources
//   -> 4x (SSIM + L1) -> +noise -> per-pixel min / argmin -> (mask) -> mean, plus edge-aware smoothness
//   (layers.py:16-25, 192-197, 211-222, 231-242, 277-290; train.py:956-1051) --
// restructured around the memory system:
//
//   * two persistent CTAs per SM walk the 32x16 tiles of all images (static round-robin), each with a two-slot pipeline: up to
//     four tiles per SM are in flight;
//   * warp 0 (producer) stages the dense inputs of a tile -- disparity, target, both sources with a 1-px halo, tie-break
//     noise, mask: everything that is NOT a data-dependent gather, 40 of the 48 algorithmic bytes per pixel -- with TMA box
//     loads (cp.async.bulk.tensor, zero fill outside the image) into a shared-memory slot behind an mbarrier;
//   * 4 gather warps complete the slot: reflection padding of the staged planes on border tiles, projection of every halo
//     position (bit-exact coordinate chain of common.cuh), 24 bilinear corner loads per position through L1, warped
//     candidates stored as float2 {warp0, warp1} for packed FFMA2;
//   * 4 SSIM warps consume the slot: separable 3x3 window sums straight from the staged planes (register ring), SSIM + L1
//     of the four candidates, min / argmin, mask, smoothness, warp-shuffle reductions, fixed-point accumulation.
// The three roles only meet at mbarriers (full -> ready -> empty per slot), so the latency-bound gathers of one tile run
// under the FP32-bound statistics of another instead of alternating with them behind __syncthreads.
// Requires W % 4 == 0 (TMA global strides are multiples of 16 bytes), W >= 40, H >= 18; other shapes take f1_fwd.cu.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "f1.cuh"
#include "tc_common.cuh"

namespace mvf {

namespace {

using namespace tc;

constexpr int TW = 32, TH = 16;
constexpr int HWD = TW + 2, HHT = TH + 2;   // tile + 1-px halo
constexpr int SWD = TW + 8;                 // staged row: [tx0 - 4, tx0 + TW + 4), 16-byte aligned start
constexpr int NPOSN = HHT * HWD;            // 612 halo positions
// Role sizes, measured at B12 192x640 (profiles/r2_f1_variants.md): the SSIM warps are the critical role and every additional gather
// warp takes issue slots from them -- 4 gather warps (5 rounds over the halo positions) 80.8 us, 5 (4 rounds) 87.3, 7 (3 rounds) 97,
// 3 (7 rounds: the gathers become critical) 99.9; twice the SSIM warps with half the rows each (MVF_F1_RPT=4) 86.
#ifndef MVF_F1_NG
#define MVF_F1_NG 4
#endif
#ifndef MVF_F1_RPT
#define MVF_F1_RPT 8
#endif
constexpr int RPT = MVF_F1_RPT;             // output rows per SSIM thread
constexpr int NG = MVF_F1_NG, NS = 2 * (16 / RPT);   // gather / SSIM warps (one SSIM thread per (pass, column, group of RPT rows))
constexpr int NSLOT = 2;                    // pipeline depth (tiles in flight per CTA)
constexpr int NGT = NG * 32, NST = NS * 32;
constexpr int NT = 32 * (1 + NG + NS);
constexpr int NROUND = (NPOSN + NGT - 1) / NGT;   // halo positions per gather thread
constexpr int CST_MAXB = 32;                // images whose matrices are kept in shared memory (others: per-tile reload)
constexpr int HOFF = 3;                     // staged column of halo column 0 (x = tx0 - 1)
static_assert(NST == 2 * TW * (TH / RPT), "one (pass, column, row group) item per SSIM thread");
static_assert(NST * 4 >= TW * TH, "four pixels per SSIM thread in phase 3");

struct __align__(128) Stage {          // TMA destinations (each 128-byte aligned), planar
    alignas(128) float disp[HHT][SWD];
    alignas(128) float tgt[3][HHT][SWD];
    alignas(128) float s0[3][HHT][SWD];
    alignas(128) float s1[3][HHT][SWD];
    alignas(128) float noise[2][TH][TW];
    alignas(128) float mask[TH][TW];
};
struct __align__(128) Smem {
    Stage st[NSLOT];
    float2 Wp[NSLOT][3][HHT][HWD];     // {warp0, warp1} at tile + halo
    float2 REPA[TH][TW];               // {id0, id1} summed over channels
    float2 REPB[TH][TW];               // {w0, w1}
    float cst[CST_MAXB][36];           // per image: inv_K rows 0..2 (12), P0 (12), P1 (12)
    uint64_t full[NSLOT], ready[NSLOT], empty[NSLOT];
    int is_last;
};

struct F1Maps {
    CUtensorMap disp, tgt, src0, src1, noise, mask;
};

struct HS {
    float t, tt;
    float2 x, xx, xt;
};

__device__ __forceinline__ float2 rep_window(const HS& a, const HS& b, const HS& c, float tc, float2 xc, float cS, float cL);

// horizontal 3-tap sums of one row for this thread's output column (staged column col + HOFF - 1 .. + 1).
// PASS 0: candidates {src0, src1} from the two staged planes; PASS 1: {warp0, warp1} from the float2 plane of the gather warps.
template <int PASS>
__device__ __forceinline__ void hsum_row(const float (*__restrict__ Tc)[SWD], const float (*__restrict__ Ac)[SWD],
                                         const float (*__restrict__ Bc)[SWD], const float2 (*__restrict__ Wc)[HWD], int hr, int col,
                                         HS& h, float& tc, float2& xc) {
    const float* tp = &Tc[hr][col + HOFF];
    const float t0 = tp[0], t1 = tp[1], t2 = tp[2];
    float2 x0, x1, x2;
    if (PASS == 0) {
        const float* ap = &Ac[hr][col + HOFF];
        const float* bp = &Bc[hr][col + HOFF];
        x0 = make_float2(ap[0], bp[0]); x1 = make_float2(ap[1], bp[1]); x2 = make_float2(ap[2], bp[2]);
    } else {
        const float2* wp = &Wc[hr][col];
        x0 = wp[0]; x1 = wp[1]; x2 = wp[2];
    }
    h.t = (t0 + t1) + t2;
    h.tt = fmaf(t2, t2, fmaf(t1, t1, t0 * t0));
    h.x = add2(add2(x0, x1), x2);
    h.xx = fma2(x2, x2, fma2(x1, x1, mul2(x0, x0)));
    h.xt = fma2(x2, f2(t2), fma2(x1, f2(t1), mul2(x0, f2(t0))));
    tc = t1;
    xc = x1;
}

// one thread: one output column, RPT rows, three channels; vertical sums from a two-row register ring
template <int PASS>
__device__ __forceinline__ void window_pass(const Stage& st, const float2 (*__restrict__ Wp)[HHT][HWD], int rg, int col, float cS,
                                            float cL, float2 acc2[RPT]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc2[r] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        HS r0, r1, cu;
        float tcp, tc;
        float2 xcp, xc;
        hsum_row<PASS>(st.tgt[c], st.s0[c], st.s1[c], Wp[c], rg * RPT + 0, col, r0, tc, xc);
        hsum_row<PASS>(st.tgt[c], st.s0[c], st.s1[c], Wp[c], rg * RPT + 1, col, r1, tcp, xcp);
#pragma unroll
        for (int r = 2; r < RPT + 2; ++r) {
            hsum_row<PASS>(st.tgt[c], st.s0[c], st.s1[c], Wp[c], rg * RPT + r, col, cu, tc, xc);
            acc2[r - 2] = add2(acc2[r - 2], rep_window(r0, r1, cu, tcp, xcp, cS, cL));   // fixed channel order
            r0 = r1;
            r1 = cu;
            tcp = tc;
            xcp = xc;
        }
    }
}

// 0.85/3 * SSIM-loss + 0.15/3 * |t - x| of one window for the packed candidate pair (layers.py:277-290, train.py:973-985)
__device__ __forceinline__ float2 rep_window(const HS& a, const HS& b, const HS& c, float tc, float2 xc, float cS, float cL) {
    const float k9 = 1.0f / 9.0f, C1 = 0.0001f, C2 = 0.0009f;
    const float vt = (a.t + b.t) + c.t, vtt = (a.tt + b.tt) + c.tt;
    const float my = vt * k9, my2 = my * my;
    const float sigyc = fmaf(vtt, k9, -my2) + C2, my2c = my2 + C1;
    const float2 vx = add2(add2(a.x, b.x), c.x), vxx = add2(add2(a.xx, b.xx), c.xx), vxt = add2(add2(a.xt, b.xt), c.xt);
    const float2 mx = mul2(vx, f2(k9));
    const float2 mxmy = mul2(mx, f2(my));
    const float2 mx2 = mul2(mx, mx);
    const float2 sigx = fma2(vxx, f2(k9), -mx2);
    const float2 sigxy = fma2(vxt, f2(k9), -mxmy);
    const float2 n = mul2(fma2(f2(2.0f), mxmy, f2(C1)), fma2(f2(2.0f), sigxy, f2(C2)));
    const float2 d = mul2(add2(mx2, f2(my2c)), add2(sigx, f2(sigyc)));
    float2 r;
    r.x = __saturatef(fmaf(-0.5f, __fdividef(n.x, d.x), 0.5f));
    r.y = __saturatef(fmaf(-0.5f, __fdividef(n.y, d.y), 0.5f));
    r.x = fmaf(cL, fabsf(tc - xc.x), cS * r.x);
    r.y = fmaf(cL, fabsf(tc - xc.y), cS * r.y);
    return r;
}

// Barrier hand-offs: ONE arrival per warp (an mbarrier arrival per thread serialises 128-160 updates of one word per tile
// and hand-off; measured 28 -> 20 us for the empty pipeline).  __syncwarp orders the other lanes' shared-memory accesses
// before the arrival.  Waiting is done by the whole warp (a single-lane spin loop starved the working warps: 96 -> 160 us).
#ifndef MVF_F1_SLEEP
#define MVF_F1_SLEEP 0
#endif
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity, int lane) {
    (void)lane;
    while (!mbar_try_wait(bar, parity)) {   // whole warp: one broadcast request; the hardware suspends the warp inside try_wait
        if (MVF_F1_SLEEP > 0) __nanosleep(MVF_F1_SLEEP);   // polling costs issue slots the working warps need
    }
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// tile index -> (image, tile origin)
struct TileXY {
    int b, tx0, ty0;
};
__device__ __forceinline__ TileXY tile_of(int t, int tiles_x, int tiles_y) {
    TileXY r;
    const int per = tiles_x * tiles_y;
    r.b = t / per;
    const int q = t - r.b * per;
    const int ty = q / tiles_x;
    r.ty0 = ty * TH;
    r.tx0 = (q - ty * tiles_x) * TW;
    return r;
}

template <bool DBG>
__global__ void __launch_bounds__(NT, 2) f1_fwd_tma_kernel(const F1Args a, const __grid_constant__ F1Maps maps, const int tiles_x,
                                                           const int tiles_y, const int n_tiles_, const int nid_loaded,
                                                           const uint32_t tx_bytes, const int dbg_mode) {
    const int n_tiles = (dbg_mode & 64) ? 0 : n_tiles_;   // timing experiment: launch + prologue + epilogue only
    // The declared alignment -- not an integer round-up of the pointer -- keeps the shared address space visible to the compiler:
    // every access below is an LDS / STS with an immediate offset.  The round-up (round 2's first version) turned all 300 of them into
    // generic 64-bit LD / ST with their own address arithmetic: 95.7 -> 87.3 us for this change alone.
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;
    const int HWi = H * W;

    if (tid == 0) {
        tma_prefetch_desc(&maps.disp);
        tma_prefetch_desc(&maps.tgt);
        tma_prefetch_desc(&maps.src0);
        tma_prefetch_desc(&maps.src1);
        if (nid_loaded) tma_prefetch_desc(&maps.noise);
        if (a.mask) tma_prefetch_desc(&maps.mask);
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.ready[s], NG);    // one arrival per gather warp (lane 0, after __syncwarp)
            mbar_init(&sm.empty[s], NS);    // one arrival per SSIM warp
        }
        fence_barrier_init();
    }
    const bool cst_resident = a.B <= CST_MAXB;
    if (cst_resident) {
        for (int i = tid; i < a.B * 36; i += NT) {
            const int b = i / 36, e = i - 36 * b;
            sm.cst[b][e] = e < 12 ? __ldg(a.inv_K + 16 * b + e) : (e < 24 ? __ldg(a.P0 + 12 * b + e - 12) : __ldg(a.P1 + 12 * b + e - 24));
        }
    }
    __syncthreads();

    if (warp == 0) {
        // ================= producer: TMA box loads of the dense inputs of each tile =================
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % NSLOT, k = it / NSLOT;
                const TileXY q = tile_of(t, tiles_x, tiles_y);
                while (!mbar_try_wait(&sm.empty[s], (k & 1) ^ 1)) __nanosleep(64);
                if (dbg_mode & 8) {   // timing experiment: no TMA traffic
                    mbar_arrive(&sm.full[s]);
                    continue;
                }
                mbar_arrive_expect_tx(&sm.full[s], tx_bytes);
                Stage& st = sm.st[s];
                tma_load_3d(&st.disp[0][0], &maps.disp, &sm.full[s], q.tx0 - 4, q.ty0 - 1, q.b);
                tma_load_3d(&st.tgt[0][0][0], &maps.tgt, &sm.full[s], q.tx0 - 4, q.ty0 - 1, 3 * q.b);
                tma_load_3d(&st.s0[0][0][0], &maps.src0, &sm.full[s], q.tx0 - 4, q.ty0 - 1, 3 * q.b);
                tma_load_3d(&st.s1[0][0][0], &maps.src1, &sm.full[s], q.tx0 - 4, q.ty0 - 1, 3 * q.b);
                if (nid_loaded) tma_load_3d(&st.noise[0][0][0], &maps.noise, &sm.full[s], q.tx0, q.ty0, nid_loaded * q.b);
                if (a.mask) tma_load_3d(&st.mask[0][0], &maps.mask, &sm.full[s], q.tx0, q.ty0, q.b);
            }
        }
    } else if (warp <= NG) {
        // ================= gather warps: reflection padding, view synthesis at tile + halo =================
        const int gt = tid - 32;
        const Geo g = make_geo(H, W);
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int s = it % NSLOT, k = it / NSLOT;
            const TileXY q = tile_of(t, tiles_x, tiles_y);
            const float* __restrict__ s0b = a.src0 + (size_t)q.b * 3 * HWi;
            const float* __restrict__ s1b = a.src1 + (size_t)q.b * 3 * HWi;
            float cv = 0.f;
            if (!cst_resident) {   // more images than the resident table holds: reload this slot's matrices
                if (gt < 12) cv = __ldg(a.inv_K + 16 * q.b + gt);
                else if (gt < 24) cv = __ldg(a.P0 + 12 * q.b + gt - 12);
                else if (gt < 36) cv = __ldg(a.P1 + 12 * q.b + gt - 24);
            }
            warp_wait(&sm.full[s], k & 1, lane);   // implies: the SSIM warps released the slot two tiles ago
            if (!cst_resident && gt < 36) sm.cst[s][gt] = cv;
            Stage& st = sm.st[s];
            // nn.ReflectionPad2d(1) of the nine staged image planes (layers.py:272): only tiles on the image border have
            // halo positions outside the image; TMA zero-filled them, the reflected column / row is inside the same box
            const bool left = q.tx0 == 0, right = q.tx0 + TW >= W, top = q.ty0 == 0, bottom = q.ty0 + TH >= H;
            if (left || right) {
                for (int i = gt; i < 9 * HHT; i += NGT) {
                    const int pl = i / HHT, r = i - pl * HHT;
                    float* row = (pl < 3 ? &st.tgt[pl][r][0] : (pl < 6 ? &st.s0[pl - 3][r][0] : &st.s1[pl - 6][r][0]));
                    if (left) row[HOFF] = row[HOFF + 2];
                    if (right) row[W - q.tx0 + HOFF + 1] = row[W - q.tx0 + HOFF - 1];
                }
                named_barrier_sync(1, NGT);
            }
            if (top || bottom) {
                for (int i = gt; i < 9 * SWD; i += NGT) {
                    const int pl = i / SWD, cidx = i - pl * SWD;
                    float (*pp)[SWD] = (pl < 3 ? st.tgt[pl] : (pl < 6 ? st.s0[pl - 3] : st.s1[pl - 6]));
                    if (top) pp[0][cidx] = pp[2][cidx];
                    if (bottom) pp[H - q.ty0 + 1][cidx] = pp[H - q.ty0 - 1][cidx];
                }
            }
            if (top || bottom || !cst_resident) named_barrier_sync(1, NGT);    // padding (and matrices) visible to every gather warp
            const float* cst = sm.cst[cst_resident ? q.b : s];
#pragma unroll 1
            for (int j = 0; j < NROUND; ++j) {
                const int p = gt + j * NGT;
                if (p >= NPOSN || (dbg_mode & 32)) break;
                const int hyj = p / HWD, hxj = p - hyj * HWD;
                const int y = q.ty0 - 1 + hyj, x = q.tx0 - 1 + hxj;
                // the warp of a padded position is the warp of its reflected pixel; rows / columns past the reflected one
                // are never used by a valid output: they are pinned to the staged box
                const int ry = max(clampi(reflect1(y, H), 0, H - 1), q.ty0 - 1);
                const int rx = max(clampi(reflect1(x, W), 0, W - 1), q.tx0 - 1);
                const float d = st.disp[ry - (q.ty0 - 1)][rx - (q.tx0 - 4)];
                Tap t0, t1;
                if (dbg_mode & 4) {   // timing experiment: no projection (samples the pixel's own position)
                    t0.x0 = rx; t0.y0 = ry; t0.fw = d; t0.fn = 0.25f;
                    t1 = t0;
                } else {
                    const float depth = disp_to_depth(d, a.min_disp, a.disp_range);
                    float cr[3], X[3], pr[3];
                    cam_ray(cst, (float)rx, (float)ry, cr);
                    project_tap(depth, cr, cst + 12, g, t0, X, pr);
                    project_tap(depth, cr, cst + 24, g, t1, X, pr);
                }
                const Corner c0 = corner_of(t0, H, W), c1 = corner_of(t1, H, W);
                const bool interior = ry == y && rx == x && hyj >= 1 && hyj <= TH && hxj >= 1 && hxj <= TW;
                if (DBG) {
                    if (a.x0y0 != nullptr && interior) {
                        const size_t n = (size_t)a.B * HWi, o = (size_t)q.b * HWi + (size_t)y * W + x;
                        a.x0y0[o] = t0.x0;
                        a.x0y0[n + o] = t0.y0;
                        a.x0y0[2 * n + o] = t1.x0;
                        a.x0y0[3 * n + o] = t1.y0;
                    }
                }
                float v0[3][4], v1[3][4];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* r0 = s0b + (c0.off + c * HWi);
                    const float* r1 = s1b + (c1.off + c * HWi);
                    if (dbg_mode & 1) {   // timing experiment: no corner loads
                        v0[c][0] = v0[c][1] = v0[c][2] = v0[c][3] = c0.fw;
                        v1[c][0] = v1[c][1] = v1[c][2] = v1[c][3] = c1.fn;
                        continue;
                    }
                    v0[c][0] = __ldg(r0); v0[c][1] = __ldg(r0 + 1); v0[c][2] = __ldg(r0 + W); v0[c][3] = __ldg(r0 + W + 1);
                    v1[c][0] = __ldg(r1); v1[c][1] = __ldg(r1 + 1); v1[c][2] = __ldg(r1 + W); v1[c][3] = __ldg(r1 + W + 1);
                }
                const float w0 = c0.fw, e0 = 1.0f - w0, n0 = c0.fn, q0 = 1.0f - n0;
                const float w1 = c1.fw, e1 = 1.0f - w1, n1 = c1.fn, q1 = 1.0f - n1;
                const float a0 = q0 * e0, b0 = q0 * w0, d0 = n0 * e0, f0 = n0 * w0;
                const float a1 = q1 * e1, b1 = q1 * w1, d1 = n1 * e1, f1 = n1 * w1;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float o0 = v0[c][0] * a0 + v0[c][1] * b0 + v0[c][2] * d0 + v0[c][3] * f0;
                    const float o1 = v1[c][0] * a1 + v1[c][1] * b1 + v1[c][2] * d1 + v1[c][3] * f1;
                    sm.Wp[s][c][hyj][hxj] = make_float2(o0, o1);
                    if (DBG) {
                        if (a.warp0 != nullptr && interior) {
                            a.warp0[((size_t)q.b * 3 + c) * HWi + (size_t)y * W + x] = o0;
                            a.warp1[((size_t)q.b * 3 + c) * HWi + (size_t)y * W + x] = o1;
                        }
                    }
                }
            }
            if (left || right || top || bottom) fence_proxy_async();   // the padding writes precede the next TMA fill of this slot
            warp_arrive(&sm.ready[s], lane);
        }
    } else {
        // ================= SSIM warps: window statistics, min / argmin, smoothness, reductions =================
        const int st_ = tid - 32 * (1 + NG);
        const bool nossim = (a.flags & F_NO_SSIM) != 0;
        const float cS = nossim ? 0.0f : 0.85f / 3.0f, cL = nossim ? 1.0f / 3.0f : 0.15f / 3.0f;
        const bool avg = (a.flags & F_AVG_REPROJECTION) != 0, am = !(a.flags & F_DISABLE_AUTOMASKING);
        const int pass = st_ / (NST / 2), rg = (st_ / 32) % (TH / RPT), col = st_ & 31;
        const int prow = (st_ >> 3) & (TH - 1), pcol = (st_ & 7) * 4;   // phase 3: 4 consecutive pixels of one row (threads 0..127)
        float photo = 0.f, sx = 0.f, sy = 0.f, sd = 0.f;
        int cur_b = -1;
        long long* acc = ws_fwd_acc(a.ws);
        auto flush = [&]() {
            if (cur_b < 0) return;
            const float v0 = warp_sum(photo), v1 = warp_sum(sx), v2 = warp_sum(sy), v3 = warp_sum(sd);
            if (lane < 4) {
                const float v = lane == 0 ? v0 : (lane == 1 ? v1 : (lane == 2 ? v2 : v3));
                atomicAdd(reinterpret_cast<unsigned long long*>(acc + 4 * cur_b + lane), (unsigned long long)to_fix((double)v));
            }
            photo = sx = sy = sd = 0.f;
        };
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int s = it % NSLOT, k = it / NSLOT;
            const TileXY q = tile_of(t, tiles_x, tiles_y);
            if (q.b != cur_b) {
                flush();
                cur_b = q.b;
            }
            warp_wait(&sm.ready[s], k & 1, lane);
            const Stage& st = sm.st[s];
            // ---- phase 2: one (pass, column, row group) item per thread, three channels summed in registers ----
            if (!(dbg_mode & 2)) {
                float2 acc2[RPT];
                if (pass == 0) window_pass<0>(st, sm.Wp[s], rg, col, cS, cL, acc2);
                else window_pass<1>(st, sm.Wp[s], rg, col, cS, cL, acc2);
                float2 (*REP)[TW] = pass == 0 ? sm.REPA : sm.REPB;
#pragma unroll
                for (int r = 0; r < RPT; ++r) REP[rg * RPT + r][col] = acc2[r];
            }
            named_barrier_sync(2, NST);
            // ---- phase 3: channel mix, min / argmin, mask, smoothness (four pixels of one row per thread) ----
            {
                const int y = q.ty0 + prow, x0 = q.tx0 + pcol;
                if (st_ * 4 < TW * TH && y < H && x0 < W && !(dbg_mode & 16)) {   // W % 4 == 0: the four pixels are all inside or all outside
                    const float4* ra = reinterpret_cast<const float4*>(&sm.REPA[prow][pcol]);
                    const float4* rb = reinterpret_cast<const float4*>(&sm.REPB[prow][pcol]);
                    const float4 A0 = ra[0], A1 = ra[1], B0 = rb[0], B1 = rb[1];
                    const float id0[4] = {A0.x, A0.z, A1.x, A1.z}, id1[4] = {A0.y, A0.w, A1.y, A1.w};
                    const float w0[4] = {B0.x, B0.z, B1.x, B1.z}, w1[4] = {B0.y, B0.w, B1.y, B1.w};
                    float nz0[4] = {0.f, 0.f, 0.f, 0.f}, nz1[4] = {0.f, 0.f, 0.f, 0.f}, mk[4] = {1.f, 1.f, 1.f, 1.f};
                    if (nid_loaded) {
                        const float4 n0 = *reinterpret_cast<const float4*>(&st.noise[0][prow][pcol]);
                        nz0[0] = n0.x; nz0[1] = n0.y; nz0[2] = n0.z; nz0[3] = n0.w;
                        if (nid_loaded > 1) {
                            const float4 n1 = *reinterpret_cast<const float4*>(&st.noise[1][prow][pcol]);
                            nz1[0] = n1.x; nz1[1] = n1.y; nz1[2] = n1.z; nz1[3] = n1.w;
                        }
                    }
                    if (a.mask) {
                        const float4 m4 = *reinterpret_cast<const float4*>(&st.mask[prow][pcol]);
                        mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
                    }
                    // disparity and target of the pixel, its right and its lower neighbour (staged planes, halo offset (1, 4))
                    const float* dr = &st.disp[prow + 1][pcol + 4];
                    const float* dn = &st.disp[prow + 2][pcol + 4];
                    const float4 dv = *reinterpret_cast<const float4*>(dr), dnv = *reinterpret_cast<const float4*>(dn);
                    const float dd[5] = {dv.x, dv.y, dv.z, dv.w, dr[4]};
                    const float db[4] = {dnv.x, dnv.y, dnv.z, dnv.w};
                    float gxs[4] = {0.f, 0.f, 0.f, 0.f}, gys[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float* tr = &st.tgt[c][prow + 1][pcol + 4];
                        const float* tn = &st.tgt[c][prow + 2][pcol + 4];
                        const float4 tv = *reinterpret_cast<const float4*>(tr), tb = *reinterpret_cast<const float4*>(tn);
                        const float tt[5] = {tv.x, tv.y, tv.z, tv.w, tr[4]};
                        const float tl[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            gxs[i] += fabsf(tt[i] - tt[i + 1]);
                            gys[i] += fabsf(tt[i] - tl[i]);
                        }
                    }
                    uint32_t packed = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int x = x0 + i;
                        float m;
                        int best = 0;
                        if (!avg) {
                            if (am) {
                                m = id0[i] + nz0[i] * 0.00001f;
                                const float c1 = id1[i] + nz1[i] * 0.00001f;
                                if (c1 < m) { m = c1; best = 1; }
                                if (w0[i] < m) { m = w0[i]; best = 2; }
                                if (w1[i] < m) { m = w1[i]; best = 3; }
                            } else {
                                m = w0[i];
                                if (w1[i] < m) { m = w1[i]; best = 1; }
                            }
                        } else {
                            const float wa = (w0[i] + w1[i]) * 0.5f;
                            if (am) {
                                m = (id0[i] + id1[i]) * 0.5f + nz0[i] * 0.00001f;
                                if (wa < m) { m = wa; best = 1; }
                            } else {
                                m = wa;
                            }
                        }
                        if (a.mask) m *= mk[i];
                        packed |= (uint32_t)best << (8 * i);
                        if (DBG) {
                            if (a.to_opt) a.to_opt[(size_t)q.b * HWi + (size_t)y * W + x] = m;
                        }
                        photo += m;
                        sd += dd[i];
                        if (x + 1 < W) sx += fabsf(dd[i] - dd[i + 1]) * __expf(-(gxs[i] / 3.0f));
                        if (y + 1 < H) sy += fabsf(dd[i] - db[i]) * __expf(-(gys[i] / 3.0f));
                    }
                    *reinterpret_cast<uint32_t*>(a.idx + (size_t)q.b * HWi + (size_t)y * W + x0) = packed;
                }
            }
            warp_arrive(&sm.empty[s], lane);
            named_barrier_sync(2, NST);    // REPA / REPB are single-buffered: everyone is done reading them
        }
        flush();
        __threadfence();
    }

    // ---------------- the last CTA turns the fixed-point sums into the outputs (and cleans them) ----------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(&a.ws->counter_fwd, 1u);
        sm.is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (sm.is_last) {
        // one thread per image (the serial version cost 14 us: 4 dependent L2 round trips + 3 double divisions per image),
        // then a fixed-order sum over images by thread 0
        __threadfence();
        long long* acc = ws_fwd_acc(a.ws);
        double* part = reinterpret_cast<double*>(&sm.st[0]);   // [3][NT] scratch (the pipeline is drained)
        const int B = a.B;
        const double HW = (double)H * (double)W;
        double ph = 0, smx = 0, smy = 0;
        for (int bb = tid; bb < B; bb += NT) {
            volatile long long* v = acc + 4 * bb;
            const long long q0 = v[0], q1 = v[1], q2 = v[2], q3 = v[3];
            v[0] = 0; v[1] = 0; v[2] = 0; v[3] = 0;
            const double Sx = from_fix(q1), Sy = from_fix(q2), Sd = from_fix(q3);
            const float mean = (float)(Sd / HW);
            const double den = (double)(mean + 1e-7f);
            ph += from_fix(q0);
            smx += Sx / den;
            smy += Sy / den;
            a.stats[4 * bb + 0] = mean;
            a.stats[4 * bb + 1] = (float)Sx;
            a.stats[4 * bb + 2] = (float)Sy;
            a.stats[4 * bb + 3] = 0.f;
        }
        part[tid] = ph;
        part[NT + tid] = smx;
        part[2 * NT + tid] = smy;
        __syncthreads();
        if (tid == 0) {
            double photo_t = 0, sx_t = 0, sy_t = 0;
            const int n_used = B < NT ? B : NT;
            for (int i = 0; i < n_used; ++i) {
                photo_t += part[i];
                sx_t += part[NT + i];
                sy_t += part[2 * NT + i];
            }
            const double n = (double)B * HW;
            const double phm = photo_t / n;
            const double smooth = sx_t / ((double)B * H * (W - 1)) + sy_t / ((double)B * (H - 1) * W);
            a.loss[0] = (float)(phm + (double)a.smooth_w * smooth);
            a.loss[1] = (float)phm;
            a.loss[2] = (float)smooth;
            a.loss[3] = 0.f;
            a.ws->counter_fwd = 0;
            __threadfence();
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn f1_encode_fn() {
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

// [planes, H, W] fp32 as a 3-D tensor map with a (bw x bh x bz) box
bool encode_planes(EncodeTiledFn enc, CUtensorMap* m, const float* base, int planes, int H, int W, int bw, int bh, int bz) {
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bz};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool f1_forward_tma_eligible(const F1Args& a) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = std::getenv("MVF_F1_TMA");
        enabled = e ? (std::atoi(e) != 0) : 1;
    }
    if (!enabled || f1_encode_fn() == nullptr) return false;
    if ((a.W & 3) != 0 || a.W < SWD || a.H < HHT) return false;
    const uintptr_t al = (uintptr_t)a.disp | (uintptr_t)a.tgt | (uintptr_t)a.src0 | (uintptr_t)a.src1 | (uintptr_t)a.noise | (uintptr_t)a.mask |
                         (uintptr_t)a.idx;
    return (al & 15) == 0;
}

cudaError_t launch_f1_forward_tma(const F1Args& a, cudaStream_t stream) {
    EncodeTiledFn enc = f1_encode_fn();
    const bool avg = (a.flags & F_AVG_REPROJECTION) != 0, am = !(a.flags & F_DISABLE_AUTOMASKING);
    const int nid = (a.noise && am) ? (avg ? 1 : 2) : 0;
    F1Maps maps;
    bool ok = encode_planes(enc, &maps.disp, a.disp, a.B, a.H, a.W, SWD, HHT, 1) &&
              encode_planes(enc, &maps.tgt, a.tgt, 3 * a.B, a.H, a.W, SWD, HHT, 3) &&
              encode_planes(enc, &maps.src0, a.src0, 3 * a.B, a.H, a.W, SWD, HHT, 3) &&
              encode_planes(enc, &maps.src1, a.src1, 3 * a.B, a.H, a.W, SWD, HHT, 3);
    maps.noise = maps.disp;
    maps.mask = maps.disp;
    if (ok && nid) ok = encode_planes(enc, &maps.noise, a.noise, nid * a.B, a.H, a.W, TW, TH, nid);
    if (ok && a.mask) ok = encode_planes(enc, &maps.mask, a.mask, a.B, a.H, a.W, TW, TH, 1);
    if (!ok) return cudaErrorInvalidValue;
    const uint32_t tx = (uint32_t)(10 * HHT * SWD * 4 + nid * TH * TW * 4 + (a.mask ? TH * TW * 4 : 0));
    const size_t smem = sizeof(Smem);
    static_assert(2 * (sizeof(Smem) + 1024) <= 228 * 1024, "two CTAs per SM");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(f1_fwd_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(f1_fwd_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    const int tiles_x = (a.W + TW - 1) / TW, tiles_y = (a.H + TH - 1) / TH, n_tiles = tiles_x * tiles_y * a.B;
    const int grid = n_tiles < 2 * n_sm ? n_tiles : 2 * n_sm;
    static int dbg = -1;   // MVF_F1_DBG: timing experiments (1 no corner loads, 2 no window statistics, 4 no projection, 8 no TMA, 16 no phase 3, 32 no gather loop) -- wrong results
    if (dbg < 0) {
        const char* e = std::getenv("MVF_F1_DBG");
        dbg = e ? std::atoi(e) : 0;
    }
    if (a.x0y0 != nullptr || a.warp0 != nullptr || a.to_opt != nullptr)
        f1_fwd_tma_kernel<true><<<grid, NT, smem, stream>>>(a, maps, tiles_x, tiles_y, n_tiles, nid, tx, dbg);
    else
        f1_fwd_tma_kernel<false><<<grid, NT, smem, stream>>>(a, maps, tiles_x, tiles_y, n_tiles, nid, tx, dbg);
    return cudaGetLastError();
}

}  // namespace mvf
