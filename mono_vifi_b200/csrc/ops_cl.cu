// Channels-last data-movement kernels of the depth decoder (HBM-bound, 16-byte vector accesses, one pass each):
//   upcat_pad : y = ReflectionPad2d(1)( cat( nearest-upsample-x2(a) or a , skip ) )       (monodepth2.py:86-93 + layers.py:126-139)
// i.e. what the reference runs as F.interpolate + torch.cat + nn.ReflectionPad2d before every 3x3 decoder convolution,
// fused into one read of the sources and one write of the padded tensor; the backward is the exact adjoint (gather form).
// All tensors are NCHW-shaped, channels-last in memory and dense: element (b,c,y,x) at ((b*H + y)*W + x)*C + c.
#include "ops_cl.cuh"
#include "pdl.cuh"

namespace mvf {
namespace {

// i -> (i0, i1, i2, i3) with i = ((i3 * n2 + i2) * n1 + i1) * n0 + i0.  32-bit arithmetic whenever the tensor allows it: the 64-bit
// div / mod chains (three of each per element) cost more instructions than the rest of these kernels.
__device__ __forceinline__ void split4(long long i, bool small, int n0, int n1, int n2, int& i0, int& i1, int& i2, int& i3) {
    if (small) {
        unsigned r = (unsigned)i;
        i0 = (int)(r % (unsigned)n0); r /= (unsigned)n0;
        i1 = (int)(r % (unsigned)n1); r /= (unsigned)n1;
        i2 = (int)(r % (unsigned)n2);
        i3 = (int)(r / (unsigned)n2);
    } else {
        long long r = i;
        i0 = (int)(r % n0); r /= n0;
        i1 = (int)(r % n1); r /= n1;
        i2 = (int)(r % n2);
        i3 = (int)(r / n2);
    }
}

__device__ __forceinline__ int reflect1i(int i, int n) {  // index map of ReflectionPad2d(1): -1 -> 1, n -> n-2
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

// one thread = one float4 (4 channels) of one padded output pixel
__global__ void upcat_pad_fwd_kernel(const float4* __restrict__ a, const float4* __restrict__ skip, float4* __restrict__ y, int B,
                                     int Ca4, int Cs4, int H, int W, int up) {
    pdl_sync();
    const int C4 = Ca4 + Cs4, Hp = H + 2, Wp = W + 2;
    const long long total = (long long)B * Hp * Wp * C4;
    const int Ha = up ? H / 2 : H, Wa = up ? W / 2 : W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c, px, py, b;
        split4(i, total < 0xffffffffLL, C4, Wp, Hp, c, px, py, b);
        const int sy = reflect1i(py - 1, H), sx = reflect1i(px - 1, W);
        float4 v;
        if (c < Ca4) {
            const int ay = up ? sy >> 1 : sy, ax = up ? sx >> 1 : sx;
            v = __ldg(a + (((long long)b * Ha + ay) * Wa + ax) * Ca4 + c);
        } else {
            v = __ldg(skip + (((long long)b * H + sy) * W + sx) * Cs4 + (c - Ca4));
        }
        y[i] = v;
    }
}

// gradient w.r.t. the unpadded concatenated tensor at (sy, sx): the padded positions that reflect onto it
__device__ __forceinline__ float4 pad_adjoint(const float4* __restrict__ gy, long long img_base, int C4, int c, int H, int W, int sy,
                                              int sx) {
    const int Wp = W + 2;
    int ys[2], xs[2], ny = 1, nx = 1;
    ys[0] = sy + 1;
    xs[0] = sx + 1;
    if (sy == 1) ys[ny++] = 0;
    if (sy == H - 2) ys[ny++] = H + 1;   // (H >= 3: at most one of the two fires for a given sy unless H == 3)
    if (sx == 1) xs[nx++] = 0;
    if (sx == W - 2) xs[nx++] = W + 1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < ny && iy < 2; ++iy)
        for (int ix = 0; ix < nx && ix < 2; ++ix) {
            const float4 g = __ldg(gy + (img_base + (long long)ys[iy] * Wp + xs[ix]) * C4 + c);
            acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
        }
    return acc;
}

// H == 3 or W == 3 would need three terms (both folds hit the middle pixel); the host rejects those sizes.
__global__ void upcat_pad_bwd_a_kernel(const float4* __restrict__ gy, float4* __restrict__ ga, int B, int Ca4, int Cs4, int H, int W,
                                       int up) {
    pdl_sync();
    const int C4 = Ca4 + Cs4, Hp = H + 2, Wp = W + 2;
    const int Ha = up ? H / 2 : H, Wa = up ? W / 2 : W;
    const long long total = (long long)B * Ha * Wa * Ca4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c, ax, ay, b;
        split4(i, total < 0xffffffffLL, Ca4, Wa, Ha, c, ax, ay, b);
        const long long img = (long long)b * Hp * Wp;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n = up ? 2 : 1;
        for (int dy = 0; dy < n; ++dy)
            for (int dx = 0; dx < n; ++dx) {
                const float4 g = pad_adjoint(gy, img, C4, c, H, W, up ? 2 * ay + dy : ay, up ? 2 * ax + dx : ax);
                acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
            }
        ga[i] = acc;
    }
}

__global__ void upcat_pad_bwd_skip_kernel(const float4* __restrict__ gy, float4* __restrict__ gs, int B, int Ca4, int Cs4, int H, int W) {
    pdl_sync();
    const int C4 = Ca4 + Cs4, Hp = H + 2, Wp = W + 2;
    const long long total = (long long)B * H * W * Cs4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c, sx, sy, b;
        split4(i, total < 0xffffffffLL, Cs4, W, H, c, sx, sy, b);
        gs[i] = pad_adjoint(gy, (long long)b * Hp * Wp, C4, Ca4 + c, H, W, sy, sx);
    }
}

}  // namespace

// 16 CTAs per SM and a grid-stride loop.  (One element per thread -- every load of the kernel in flight at once -- was measured: the
// step got 0.07 ms SLOWER, the large grids crowd out the kernels of the concurrent streams.)
static int grid_for(long long total) {
    long long b = (total + 255) / 256;
    return (int)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

cudaError_t upcat_pad_fwd(const float* a, const float* skip, float* y, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st) {
    const long long total = (long long)B * (H + 2) * (W + 2) * ((Ca + Cs) / 4);
    launch_pdl(upcat_pad_fwd_kernel, dim3((unsigned)(grid_for(total))), dim3(256), (size_t)(0), st, (const float4*)a, (const float4*)skip, (float4*)y, B, Ca / 4, Cs / 4, H, W, up);
    return cudaGetLastError();
}

cudaError_t upcat_pad_bwd(const float* gy, float* ga, float* gskip, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st) {
    if (ga) {
        const long long total = (long long)B * (up ? H / 2 : H) * (up ? W / 2 : W) * (Ca / 4);
        launch_pdl(upcat_pad_bwd_a_kernel, dim3((unsigned)(grid_for(total))), dim3(256), (size_t)(0), st, (const float4*)gy, (float4*)ga, B, Ca / 4, Cs / 4, H, W, up);
    }
    if (gskip && Cs > 0) {
        const long long total = (long long)B * H * W * (Cs / 4);
        launch_pdl(upcat_pad_bwd_skip_kernel, dim3((unsigned)(grid_for(total))), dim3(256), (size_t)(0), st, (const float4*)gy, (float4*)gskip, B, Ca / 4, Cs / 4, H, W);
    }
    return cudaGetLastError();
}

}  // namespace mvf

// ---- MaxPool2d(kernel 3, stride 2, padding 1), channels-last (torchvision ResNet stem: monodepth2.py:39, posenet.py:91) ----
// forward writes the pooled value and the position of the maximum inside the 3x3 window (0..8, first maximum wins, as
// ATen does); backward is a gather: every input pixel collects the gradient of the (at most four) windows whose maximum
// it is, so no atomics and a deterministic result.
namespace mvf {
namespace {

__global__ void maxpool3s2_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, uchar4* __restrict__ idx, int B, int C4,
                                      int H, int W, int Ho, int Wo) {
    pdl_sync();
    const long long total = (long long)B * Ho * Wo * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c, ox, oy, b;
        split4(i, total < 0xffffffffLL, C4, Wo, Ho, c, ox, oy, b);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uchar4 k = make_uchar4(0, 0, 0, 0);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * oy - 1 + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * ox - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                const float4 v = __ldg(x + (((long long)b * H + iy) * W + ix) * C4 + c);
                const unsigned char t = (unsigned char)(dy * 3 + dx);
                if (v.x > m.x || v.x != v.x) { m.x = v.x; k.x = t; }
                if (v.y > m.y || v.y != v.y) { m.y = v.y; k.y = t; }
                if (v.z > m.z || v.z != v.z) { m.z = v.z; k.z = t; }
                if (v.w > m.w || v.w != v.w) { m.w = v.w; k.w = t; }
            }
        }
        y[i] = m;
        idx[i] = k;
    }
}

__global__ void maxpool3s2_bwd_kernel(const float4* __restrict__ gy, const uchar4* __restrict__ idx, float4* __restrict__ gx, int B,
                                      int C4, int H, int W, int Ho, int Wo) {
    pdl_sync();
    const long long total = (long long)B * H * W * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c, ix, iy, b;
        split4(i, total < 0xffffffffLL, C4, W, H, c, ix, iy, b);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // windows (oy, ox) with 2*oy - 1 + dy == iy, dy in 0..2
        for (int oy = (iy + 1) / 2 - ((iy & 1) ? 1 : 0); oy <= (iy + 1) / 2; ++oy) {
            if (oy < 0 || oy >= Ho) continue;
            const int dy = iy - (2 * oy - 1);
            if (dy < 0 || dy > 2) continue;
            for (int ox = (ix + 1) / 2 - ((ix & 1) ? 1 : 0); ox <= (ix + 1) / 2; ++ox) {
                if (ox < 0 || ox >= Wo) continue;
                const int dx = ix - (2 * ox - 1);
                if (dx < 0 || dx > 2) continue;
                const long long o = (((long long)b * Ho + oy) * Wo + ox) * C4 + c;
                const uchar4 k = __ldg(idx + o);
                const float4 g = __ldg(gy + o);
                const unsigned char t = (unsigned char)(dy * 3 + dx);
                if (k.x == t) acc.x += g.x;
                if (k.y == t) acc.y += g.y;
                if (k.z == t) acc.z += g.z;
                if (k.w == t) acc.w += g.w;
            }
        }
        gx[i] = acc;
    }
}

}  // namespace

cudaError_t maxpool3s2_fwd(const float* x, float* y, unsigned char* idx, int B, int C, int H, int W, cudaStream_t st) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)B * Ho * Wo * (C / 4);
    launch_pdl(maxpool3s2_fwd_kernel, dim3((unsigned)(grid_for(total))), dim3(256), (size_t)(0), st, (const float4*)x, (float4*)y, (uchar4*)idx, B, C / 4, H, W, Ho, Wo);
    return cudaGetLastError();
}

cudaError_t maxpool3s2_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int C, int H, int W, cudaStream_t st) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)B * H * W * (C / 4);
    launch_pdl(maxpool3s2_bwd_kernel, dim3((unsigned)(grid_for(total))), dim3(256), (size_t)(0), st, (const float4*)gy, (const uchar4*)idx, (float4*)gx, B, C / 4, H, W, Ho, Wo);
    return cudaGetLastError();
}

}  // namespace mvf
