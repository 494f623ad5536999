/*
 * oracle/f1_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * Plain-C, single-threaded CPU restatement of the Mono-ViFI view-synthesis + photometric-loss
 * chain, used as the checker for the sm_100a kernels in
 * mono_vifi_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here against golden vectors
 * produced by importing the UNMODIFIED reference (/root/reference, torch CPU) with
 * tests/golden/gen_golden.py.  Coordinate/index outputs are compared bit-exactly.
 *
 * Reference lines restated (relative to /root/reference):
 *   disp_to_depth ................. layers.py:16-25
 *   BackprojectDepth.forward ...... layers.py:192-197
 *   Project3D.forward ............. layers.py:211-222
 *   F.grid_sample(border, align_corners=True) call ... train.py:966-969 (ATen GridSampler semantics)
 *   SSIM.forward .................. layers.py:277-290
 *   get_smooth_loss ............... layers.py:231-242
 *   compute_reprojection_loss ..... train.py:973-985
 *   compute_losses_base ........... train.py:987-1051
 *   compute_SI_log_depth_loss ..... train.py:924-941
 *
 * Rounding: on the coordinate path every op is a separately rounded fp32 op, and the small
 * matmuls are k-ordered fmaf chains (what torch-CPU's bmm produces for K=3/K=4).  Compile with
 * -ffp-contract=off so the compiler never fuses anything that is not written as fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MVFO_NO_SSIM 1
#define MVFO_AVG_REPROJECTION 2
#define MVFO_DISABLE_AUTOMASKING 4

typedef struct {
    int B, H, W;
    float min_disp;   /* (float)(1/max_depth)                  layers.py:21 */
    float disp_range; /* (float)(1/min_depth - 1/max_depth)    layers.py:22-23 */
    float smooth_w;   /* opt.disparity_smoothness              train.py:1049 */
    int flags;
} mvfo_params;

static inline int reflect1(int i, int n) { /* nn.ReflectionPad2d(1): -1 -> 1, n -> n-2 */
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

/* layers.py:16-25 */
void mvfo_disp_to_depth(const float *disp, float *scaled_disp, float *depth, long n, float min_disp,
                        float disp_range) {
    for (long i = 0; i < n; ++i) {
        float sd = min_disp + disp_range * disp[i];
        if (scaled_disp) scaled_disp[i] = sd;
        if (depth) depth[i] = 1.0f / sd;
    }
}

/* layers.py:192-197.  inv_K is [B,4,4]; out is [B,4,H*W]. */
void mvfo_backproject(const float *depth, const float *inv_K, float *out, int B, int H, int W) {
    long HW = (long)H * W;
    for (int b = 0; b < B; ++b) {
        const float *k = inv_K + 16 * b;
        for (int v = 0; v < H; ++v)
            for (int u = 0; u < W; ++u) {
                long i = (long)v * W + u;
                float d = depth[b * HW + i];
                for (int r = 0; r < 3; ++r) {
                    float acc = k[4 * r + 0] * (float)u;
                    acc = fmaf(k[4 * r + 1], (float)v, acc);
                    acc = fmaf(k[4 * r + 2], 1.0f, acc);
                    out[((long)b * 4 + r) * HW + i] = d * acc;
                }
                out[((long)b * 4 + 3) * HW + i] = 1.0f;
            }
    }
}

/* layers.py:211-222 with P = (K @ T)[:, :3, :] supplied by the caller ([B,3,4]); out is the
 * normalised sampling grid [B,H,W,2]. */
void mvfo_project(const float *points, const float *P, float *grid, int B, int H, int W) {
    long HW = (long)H * W;
    float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    for (int b = 0; b < B; ++b) {
        const float *p = P + 12 * b;
        for (long i = 0; i < HW; ++i) {
            float X[4];
            for (int j = 0; j < 4; ++j) X[j] = points[((long)b * 4 + j) * HW + i];
            float c[3];
            for (int r = 0; r < 3; ++r) {
                float acc = p[4 * r + 0] * X[0];
                acc = fmaf(p[4 * r + 1], X[1], acc);
                acc = fmaf(p[4 * r + 2], X[2], acc);
                acc = fmaf(p[4 * r + 3], X[3], acc);
                c[r] = acc;
            }
            float z = c[2] + 1e-7f;
            float x = c[0] / z, y = c[1] / z;
            x = x / wm1;
            y = y / hm1;
            grid[((long)b * HW + i) * 2 + 0] = (x - 0.5f) * 2.0f;
            grid[((long)b * HW + i) * 2 + 1] = (y - 0.5f) * 2.0f;
        }
    }
}

typedef struct {
    float ixr, iyr; /* un-normalised, before the border clip */
    float ix, iy;   /* clipped */
    int x0, y0;
    float fw, fn;   /* frac to west / north: ix-x0, iy-y0 */
} mvfo_tap;

static inline mvfo_tap tap_from_grid(float gx, float gy, int H, int W) {
    mvfo_tap t;
    t.ixr = (gx + 1.0f) * ((float)(W - 1) / 2.0f);
    t.iyr = (gy + 1.0f) * ((float)(H - 1) / 2.0f);
    t.ix = fminf((float)(W - 1), fmaxf(t.ixr, 0.0f));
    t.iy = fminf((float)(H - 1), fmaxf(t.iyr, 0.0f));
    float fx = floorf(t.ix), fy = floorf(t.iy);
    t.x0 = (int)fx;
    t.y0 = (int)fy;
    t.fw = t.ix - fx;
    t.fn = t.iy - fy;
    return t;
}

/* ATen grid_sampler_2d, bilinear, padding_mode=border, align_corners=True (train.py:966-969).
 * img [B,C,H,W] (same H,W as the grid), grid [B,H,W,2]; out [B,C,H,W]; x0/y0 optional int32 [B,H,W]. */
void mvfo_grid_sample(const float *img, const float *grid, float *out, int32_t *x0o, int32_t *y0o, int B,
                      int C, int H, int W) {
    long HW = (long)H * W;
    for (int b = 0; b < B; ++b)
        for (long i = 0; i < HW; ++i) {
            mvfo_tap t = tap_from_grid(grid[((long)b * HW + i) * 2], grid[((long)b * HW + i) * 2 + 1], H, W);
            if (x0o) x0o[b * HW + i] = t.x0;
            if (y0o) y0o[b * HW + i] = t.y0;
            int x1 = t.x0 + 1, y1 = t.y0 + 1;
            float w = t.fw, e = 1.0f - w, n = t.fn, s = 1.0f - n;
            float cnw = s * e, cne = s * w, csw = n * e, cse = n * w;
            int ex = x1 < W, sy = y1 < H;
            if (!out) continue;
            for (int c = 0; c < C; ++c) {
                const float *im = img + ((long)b * C + c) * HW;
                float acc = im[(long)t.y0 * W + t.x0] * cnw;
                if (ex) acc += im[(long)t.y0 * W + x1] * cne;
                if (sy) acc += im[(long)y1 * W + t.x0] * csw;
                if (ex && sy) acc += im[(long)y1 * W + x1] * cse;
                out[((long)b * C + c) * HW + i] = acc;
            }
        }
}

/* layers.py:277-290; x,y,out are [N,H,W] planes (N = B*C).  Window statistics are accumulated in
 * double: sigma = E[x^2]-mu^2 cancels ~3 digits, so the fp32 reference itself carries ~5e-5 absolute
 * error; the oracle is the accurate side of the comparison. */
void mvfo_ssim(const float *x, const float *y, float *out, int N, int H, int W) {
    const double C1 = (float)(0.01 * 0.01), C2 = (float)(0.03 * 0.03);
    long HW = (long)H * W;
    for (int p = 0; p < N; ++p) {
        const float *xp = x + p * HW, *yp = y + p * HW;
        for (int v = 0; v < H; ++v)
            for (int u = 0; u < W; ++u) {
                double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
                for (int dv = -1; dv <= 1; ++dv)
                    for (int du = -1; du <= 1; ++du) {
                        long j = (long)reflect1(v + dv, H) * W + reflect1(u + du, W);
                        double a = xp[j], t = yp[j];
                        sx += a;
                        sy += t;
                        sxx += a * a;
                        syy += t * t;
                        sxy += a * t;
                    }
                double mx = sx / 9.0, my = sy / 9.0;
                double sgx = sxx / 9.0 - mx * mx, sgy = syy / 9.0 - my * my, sgxy = sxy / 9.0 - mx * my;
                double n = (2.0 * mx * my + C1) * (2.0 * sgxy + C2);
                double d = (mx * mx + my * my + C1) * (sgx + sgy + C2);
                double r = (1.0 - n / d) / 2.0;
                out[p * HW + (long)v * W + u] = (float)fmin(1.0, fmax(0.0, r));
            }
    }
}

/* layers.py:231-242; disp [B,1,H,W], img [B,3,H,W]. */
double mvfo_smooth_loss(const float *disp, const float *img, int B, int H, int W) {
    long HW = (long)H * W;
    double sx = 0, sy = 0;
    for (int b = 0; b < B; ++b) {
        const float *d = disp + b * HW, *im = img + (long)b * 3 * HW;
        for (int v = 0; v < H; ++v)
            for (int u = 0; u < W; ++u) {
                long i = (long)v * W + u;
                if (u + 1 < W) {
                    float gi = 0;
                    for (int c = 0; c < 3; ++c) gi += fabsf(im[c * HW + i] - im[c * HW + i + 1]);
                    sx += fabsf(d[i] - d[i + 1]) * expf(-(gi / 3.0f));
                }
                if (v + 1 < H) {
                    float gi = 0;
                    for (int c = 0; c < 3; ++c) gi += fabsf(im[c * HW + i] - im[c * HW + i + W]);
                    sy += fabsf(d[i] - d[i + W]) * expf(-(gi / 3.0f));
                }
            }
    }
    return sx / ((double)B * H * (W - 1)) + sy / ((double)B * (H - 1) * W);
}

/* train.py:924-941; pred/target [B,1,H,W], mask [B,1,H,W] or NULL. */
double mvfo_si_log_loss(const float *pred, const float *target, const float *mask, int B, int H, int W,
                        float beta) {
    long HW = (long)H * W;
    double total = 0;
    for (int b = 0; b < B; ++b) {
        double n = 0, s1 = 0, s2 = 0;
        for (long i = 0; i < HW; ++i) {
            float m = mask ? mask[b * HW + i] : 1.0f;
            float ld = logf(pred[b * HW + i] + 1e-7f) * m - logf(target[b * HW + i] + 1e-7f) * m;
            n += m;
            s1 += ld;
            s2 += (double)ld * ld;
        }
        n += 1e-8;
        total += s2 / n - beta * s1 * s1 / (n * n);
    }
    return total / B;
}

/* d SI-log / d pred and d target (both get gradients in the reference: depth_single and depth_fused). */
void mvfo_si_log_loss_bwd(const float *pred, const float *target, const float *mask, int B, int H, int W,
                          float beta, float gout, float *g_pred, float *g_target) {
    long HW = (long)H * W;
    for (int b = 0; b < B; ++b) {
        double n = 0, s1 = 0;
        for (long i = 0; i < HW; ++i) {
            float m = mask ? mask[b * HW + i] : 1.0f;
            n += m;
            s1 += logf(pred[b * HW + i] + 1e-7f) * m - logf(target[b * HW + i] + 1e-7f) * m;
        }
        n += 1e-8;
        for (long i = 0; i < HW; ++i) {
            float m = mask ? mask[b * HW + i] : 1.0f;
            double ld = logf(pred[b * HW + i] + 1e-7f) * m - logf(target[b * HW + i] + 1e-7f) * m;
            double g = (double)gout / B * (2.0 * ld / n - 2.0 * beta * s1 / (n * n)) * m;
            if (g_pred) g_pred[b * HW + i] = (float)(g / ((double)pred[b * HW + i] + 1e-7));
            if (g_target) g_target[b * HW + i] = (float)(-g / ((double)target[b * HW + i] + 1e-7));
        }
    }
}

/* -------------------------------------------------------------------------------------------
 * Fused chain: generate_images_pred x2 + compute_losses_base  (train.py:956-1051)
 * ----------------------------------------------------------------------------------------- */

/* grid for one source from disp (composition of the three functions above, same op order) */
static void f1_grid(const mvfo_params *p, const float *disp, const float *inv_K, const float *P, float *grid) {
    long n = (long)p->B * p->H * p->W;
    float *depth = (float *)malloc(n * sizeof(float));
    float *pts = (float *)malloc(4 * n * sizeof(float));
    mvfo_disp_to_depth(disp, NULL, depth, n, p->min_disp, p->disp_range);
    mvfo_backproject(depth, inv_K, pts, p->B, p->H, p->W);
    mvfo_project(pts, P, grid, p->B, p->H, p->W);
    free(depth);
    free(pts);
}

/* train.py:973-985 : rep[B,H,W] */
static void f1_reprojection(const mvfo_params *p, const float *pred, const float *tgt, float *rep) {
    int B = p->B, H = p->H, W = p->W;
    long HW = (long)H * W;
    float *ss = NULL;
    if (!(p->flags & MVFO_NO_SSIM)) {
        ss = (float *)malloc(3 * B * HW * sizeof(float));
        mvfo_ssim(pred, tgt, ss, 3 * B, H, W);
    }
    for (int b = 0; b < B; ++b)
        for (long i = 0; i < HW; ++i) {
            float l1 = 0, sm = 0;
            for (int c = 0; c < 3; ++c) {
                long j = ((long)b * 3 + c) * HW + i;
                l1 += fabsf(tgt[j] - pred[j]);
                if (ss) sm += ss[j];
            }
            l1 /= 3.0f;
            rep[b * HW + i] = ss ? 0.85f * (sm / 3.0f) + 0.15f * l1 : l1;
        }
    free(ss);
}

/*
 * Forward.  All image tensors [B,3,H,W]; disp [B,1,H,W]; inv_K [B,4,4]; P0,P1 [B,3,4];
 * noise [B,nid,H,W] (nid = 2, or 1 with avg_reprojection; may be NULL = zero noise);
 * mask_rec [B,1,H,W] or NULL.
 * Outputs (each may be NULL): loss[3] = {total, photometric mean, smooth (unweighted)},
 *   to_optimise [B,H,W], idx uint8 [B,H,W], grid0/grid1 [B,H,W,2], warp0/warp1 [B,3,H,W],
 *   x0y0 int32 [2 sources][2 (x,y)][B,H,W].
 */
void mvfo_f1_forward(const mvfo_params *p, const float *disp, const float *tgt, const float *src0,
                     const float *src1, const float *inv_K, const float *P0, const float *P1,
                     const float *noise, const float *mask_rec, double *loss, float *to_opt, uint8_t *idx,
                     float *grid0, float *grid1, float *warp0, float *warp1, int32_t *x0y0) {
    int B = p->B, H = p->H, W = p->W;
    long HW = (long)H * W, n = B * HW;
    const float *srcs[2] = {src0, src1};
    const float *Ps[2] = {P0, P1};
    float *grids[2], *warps[2], *rep[2], *idl[2];
    for (int k = 0; k < 2; ++k) {
        grids[k] = (float *)malloc(2 * n * sizeof(float));
        warps[k] = (float *)malloc(3 * n * sizeof(float));
        rep[k] = (float *)malloc(n * sizeof(float));
        idl[k] = (float *)malloc(n * sizeof(float));
        f1_grid(p, disp, inv_K, Ps[k], grids[k]);
        mvfo_grid_sample(srcs[k], grids[k], warps[k], x0y0 ? x0y0 + (2 * k) * n : NULL,
                         x0y0 ? x0y0 + (2 * k + 1) * n : NULL, B, 3, H, W);
        f1_reprojection(p, warps[k], tgt, rep[k]);
        if (!(p->flags & MVFO_DISABLE_AUTOMASKING)) f1_reprojection(p, srcs[k], tgt, idl[k]);
    }
    int avg = (p->flags & MVFO_AVG_REPROJECTION) != 0, am = !(p->flags & MVFO_DISABLE_AUTOMASKING);
    int nid = am ? (avg ? 1 : 2) : 0;
    double photo = 0;
    for (int b = 0; b < B; ++b)
        for (long i = 0; i < HW; ++i) {
            float comb[4];
            int nc = 0;
            if (am) {
                if (avg) {
                    float v = (idl[0][b * HW + i] + idl[1][b * HW + i]) / 2.0f;
                    comb[nc++] = v + (noise ? noise[(long)b * HW + i] * 0.00001f : 0.0f);
                } else
                    for (int k = 0; k < 2; ++k)
                        comb[nc++] = idl[k][b * HW + i] +
                                     (noise ? noise[((long)b * 2 + k) * HW + i] * 0.00001f : 0.0f);
            }
            if (avg)
                comb[nc++] = (rep[0][b * HW + i] + rep[1][b * HW + i]) / 2.0f;
            else
                for (int k = 0; k < 2; ++k) comb[nc++] = rep[k][b * HW + i];
            int best = 0;
            for (int c = 1; c < nc; ++c)
                if (comb[c] < comb[best]) best = c;
            float m = comb[best];
            if (mask_rec) m *= mask_rec[b * HW + i];
            if (to_opt) to_opt[b * HW + i] = m;
            if (idx) idx[b * HW + i] = (uint8_t)best;
            photo += m;
        }
    (void)nid;
    photo /= (double)n;
    /* train.py:1044-1049 */
    float *nd = (float *)malloc(n * sizeof(float));
    for (int b = 0; b < B; ++b) {
        double s = 0;
        for (long i = 0; i < HW; ++i) s += disp[b * HW + i];
        float mean = (float)(s / HW);
        for (long i = 0; i < HW; ++i) nd[b * HW + i] = disp[b * HW + i] / (mean + 1e-7f);
    }
    double smooth = mvfo_smooth_loss(nd, tgt, B, H, W);
    free(nd);
    if (loss) {
        loss[0] = photo + (double)p->smooth_w * smooth;
        loss[1] = photo;
        loss[2] = smooth;
    }
    for (int k = 0; k < 2; ++k) {
        if (k == 0 ? grid0 != NULL : grid1 != NULL) memcpy(k == 0 ? grid0 : grid1, grids[k], 2 * n * sizeof(float));
        if (k == 0 ? warp0 != NULL : warp1 != NULL) memcpy(k == 0 ? warp0 : warp1, warps[k], 3 * n * sizeof(float));
        free(grids[k]);
        free(warps[k]);
        free(rep[k]);
        free(idl[k]);
    }
}

/*
 * Backward of mvfo_f1_forward w.r.t. disp and P0/P1 (the only inputs with gradient flow: SURVEY §9.3).
 * idx is the argmin map produced by the forward.  gout = dL/d(loss).  Analytic, accumulated in double.
 * Outputs: g_disp [B,1,H,W], g_P0/g_P1 [B,3,4].
 */
void mvfo_f1_backward(const mvfo_params *p, const float *disp, const float *tgt, const float *src0,
                      const float *src1, const float *inv_K, const float *P0, const float *P1,
                      const float *mask_rec, const uint8_t *idx, float gout, float *g_disp, float *g_P0,
                      float *g_P1) {
    int B = p->B, H = p->H, W = p->W;
    long HW = (long)H * W, n = B * HW;
    const double C1 = (float)(0.01 * 0.01), C2 = (float)(0.03 * 0.03);
    const float *srcs[2] = {src0, src1};
    const float *Ps[2] = {P0, P1};
    float *gPs[2] = {g_P0, g_P1};
    int avg = (p->flags & MVFO_AVG_REPROJECTION) != 0, am = !(p->flags & MVFO_DISABLE_AUTOMASKING);
    int nossim = (p->flags & MVFO_NO_SSIM) != 0;
    int first_rep = am ? (avg ? 1 : 2) : 0; /* channel index of the first warped entry in `combined` */
    double *gd = (double *)calloc(n, sizeof(double));
    float *grid = (float *)malloc(2 * n * sizeof(float));
    float *warp = (float *)malloc(3 * n * sizeof(float));
    double *ga = (double *)malloc(3 * n * sizeof(double));

    for (int k = 0; k < 2; ++k) {
        f1_grid(p, disp, inv_K, Ps[k], grid);
        mvfo_grid_sample(srcs[k], grid, warp, NULL, NULL, B, 3, H, W);
        memset(ga, 0, 3 * n * sizeof(double));
        /* d loss / d warped (scatter form over SSIM windows) */
        for (int b = 0; b < B; ++b)
            for (int v = 0; v < H; ++v)
                for (int u = 0; u < W; ++u) {
                    long i = (long)v * W + u;
                    double share;
                    int sel = idx ? idx[b * HW + i] : 0;
                    if (avg)
                        share = (sel == first_rep) ? 0.5 : 0.0;
                    else
                        share = (sel == first_rep + k) ? 1.0 : 0.0;
                    if (!am && avg) share = 0.5;
                    if (share == 0.0) continue;
                    double gs = (double)gout * share / (double)n * (mask_rec ? mask_rec[b * HW + i] : 1.0f);
                    for (int c = 0; c < 3; ++c) {
                        const float *a = warp + ((long)b * 3 + c) * HW, *t = tgt + ((long)b * 3 + c) * HW;
                        double *g = ga + ((long)b * 3 + c) * HW;
                        double df = (double)t[i] - a[i];
                        double sg = df > 0 ? -1.0 : (df < 0 ? 1.0 : 0.0);
                        g[i] += gs * (nossim ? 1.0 : 0.15) / 3.0 * sg;
                        if (nossim) continue;
                        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
                        long js[9];
                        int q = 0;
                        for (int dv = -1; dv <= 1; ++dv)
                            for (int du = -1; du <= 1; ++du) {
                                long j = (long)reflect1(v + dv, H) * W + reflect1(u + du, W);
                                js[q++] = j;
                                sx += a[j];
                                sy += t[j];
                                sxx += (double)a[j] * a[j];
                                syy += (double)t[j] * t[j];
                                sxy += (double)a[j] * t[j];
                            }
                        double mx = sx / 9, my = sy / 9;
                        double vx = sxx / 9 - mx * mx, vy = syy / 9 - my * my, vxy = sxy / 9 - mx * my;
                        double A = 2 * mx * my + C1, Bn = 2 * vxy + C2, Cd = mx * mx + my * my + C1,
                               D = vx + vy + C2;
                        double S = A * Bn / (Cd * D);
                        double raw = (1 - S) / 2;
                        if (raw < 0 || raw > 1) continue; /* clamp passes gradient on [0,1] */
                        double dS = gs * 0.85 / 3.0 * (-0.5);
                        double g_mu = (2 * my * Bn + A * (-2 * my)) / (Cd * D) - S * (2 * mx) / Cd - S * (-2 * mx) / D;
                        double g_xx = -S / D;
                        double g_xy = 2 * A / (Cd * D);
                        for (q = 0; q < 9; ++q) {
                            long j = js[q];
                            g[j] += dS * (g_mu + 2.0 * a[j] * g_xx + (double)t[j] * g_xy) / 9.0;
                        }
                    }
                }
        /* warped -> grid -> camera point -> depth -> disp, and -> P */
        double gP[12];
        for (int b = 0; b < B; ++b) {
            memset(gP, 0, sizeof(gP));
            const float *kk = inv_K + 16 * b, *pp = Ps[k] + 12 * b;
            for (int v = 0; v < H; ++v)
                for (int u = 0; u < W; ++u) {
                    long i = (long)v * W + u;
                    mvfo_tap t = tap_from_grid(grid[((long)b * HW + i) * 2], grid[((long)b * HW + i) * 2 + 1], H, W);
                    int x1 = t.x0 + 1, y1 = t.y0 + 1;
                    int ex = x1 < W, sy = y1 < H;
                    double w = t.fw, e = 1.0 - w, nn = t.fn, s = 1.0 - nn;
                    double gix = 0, giy = 0;
                    for (int c = 0; c < 3; ++c) {
                        const float *im = srcs[k] + ((long)b * 3 + c) * HW;
                        double vnw = im[(long)t.y0 * W + t.x0];
                        double vne = ex ? im[(long)t.y0 * W + x1] : 0.0;
                        double vsw = sy ? im[(long)y1 * W + t.x0] : 0.0;
                        double vse = (ex && sy) ? im[(long)y1 * W + x1] : 0.0;
                        double g = ga[((long)b * 3 + c) * HW + i];
                        gix += g * (s * (vne - vnw) + nn * (vse - vsw));
                        giy += g * (e * (vsw - vnw) + w * (vse - vne));
                    }
                    if (!(t.ixr > 0.0f && t.ixr < (float)(W - 1))) gix = 0;
                    if (!(t.iyr > 0.0f && t.iyr < (float)(H - 1))) giy = 0;
                    if (gix == 0 && giy == 0) continue;
                    /* recompute the geometry (same op order as the forward) */
                    float sd = p->min_disp + p->disp_range * disp[b * HW + i];
                    float depth = 1.0f / sd;
                    float cr[3], X[4], pr[3];
                    for (int r = 0; r < 3; ++r) {
                        float acc = kk[4 * r + 0] * (float)u;
                        acc = fmaf(kk[4 * r + 1], (float)v, acc);
                        acc = fmaf(kk[4 * r + 2], 1.0f, acc);
                        cr[r] = acc;
                        X[r] = depth * acc;
                    }
                    X[3] = 1.0f;
                    for (int r = 0; r < 3; ++r) {
                        float acc = pp[4 * r + 0] * X[0];
                        acc = fmaf(pp[4 * r + 1], X[1], acc);
                        acc = fmaf(pp[4 * r + 2], X[2], acc);
                        acc = fmaf(pp[4 * r + 3], X[3], acc);
                        pr[r] = acc;
                    }
                    double z = (double)(pr[2] + 1e-7f);
                    double gp[3];
                    gp[0] = gix / z;
                    gp[1] = giy / z;
                    gp[2] = -(gix * pr[0] + giy * pr[1]) / (z * z);
                    double gdepth = 0;
                    for (int r = 0; r < 3; ++r)
                        for (int j = 0; j < 4; ++j) gP[4 * r + j] += gp[r] * X[j];
                    for (int j = 0; j < 3; ++j) {
                        double gX = gp[0] * pp[j] + gp[1] * pp[4 + j] + gp[2] * pp[8 + j];
                        gdepth += gX * cr[j];
                    }
                    gd[b * HW + i] += -gdepth * (double)depth * depth * p->disp_range;
                }
            if (gPs[k])
                for (int j = 0; j < 12; ++j) gPs[k][12 * b + j] = (float)gP[j];
        }
    }
    /* smoothness term (train.py:1044-1049, layers.py:231-242) */
    for (int b = 0; b < B; ++b) {
        const float *d = disp + b * HW, *im = tgt + (long)b * 3 * HW;
        double s = 0;
        for (long i = 0; i < HW; ++i) s += d[i];
        float mean = (float)(s / HW);
        double den = (double)(mean + 1e-7f);
        double *gn = (double *)calloc(HW, sizeof(double));
        double cx = (double)gout * p->smooth_w / ((double)B * H * (W - 1));
        double cy = (double)gout * p->smooth_w / ((double)B * (H - 1) * W);
        for (int v = 0; v < H; ++v)
            for (int u = 0; u < W; ++u) {
                long i = (long)v * W + u;
                if (u + 1 < W) {
                    float gi = 0;
                    for (int c = 0; c < 3; ++c) gi += fabsf(im[c * HW + i] - im[c * HW + i + 1]);
                    double wgt = exp(-(double)(gi / 3.0f));
                    float df = d[i] / (mean + 1e-7f) - d[i + 1] / (mean + 1e-7f);
                    double sg = df > 0 ? 1.0 : (df < 0 ? -1.0 : 0.0);
                    gn[i] += cx * sg * wgt;
                    gn[i + 1] -= cx * sg * wgt;
                }
                if (v + 1 < H) {
                    float gi = 0;
                    for (int c = 0; c < 3; ++c) gi += fabsf(im[c * HW + i] - im[c * HW + i + W]);
                    double wgt = exp(-(double)(gi / 3.0f));
                    float df = d[i] / (mean + 1e-7f) - d[i + W] / (mean + 1e-7f);
                    double sg = df > 0 ? 1.0 : (df < 0 ? -1.0 : 0.0);
                    gn[i] += cy * sg * wgt;
                    gn[i + W] -= cy * sg * wgt;
                }
            }
        double dot = 0;
        for (long i = 0; i < HW; ++i) dot += gn[i] * d[i];
        for (long i = 0; i < HW; ++i) gd[b * HW + i] += gn[i] / den - dot / (den * den * (double)HW);
        free(gn);
    }
    for (long i = 0; i < n; ++i) g_disp[i] = (float)gd[i];
    free(gd);
    free(grid);
    free(warp);
    free(ga);
}

int mvfo_version(void) { return 1; }
