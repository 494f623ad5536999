// Channels-last data-movement kernels of the depth decoder (HBM-bound, 16-byte vector accesses, one pass each):
//   upcat_pad : y = ReflectionPad2d(1)( cat( nearest-upsample-x2(a) or a , skip ) )       (monodepth2.py:86-93 + layers.py:126-139)
// i.e. what the reference runs as F.interpolate + torch.cat + nn.ReflectionPad2d before every 3x3 decoder convolution,
// fused into one read of the sources and one write of the padded tensor; the backward is the exact adjoint (gather form).
// All tensors are NCHW-shaped, channels-last in memory and dense: element (b,c,y,x) at ((b*H + y)*W + x)*C + c.
#include "ops_cl.cuh"

namespace mvf {
namespace {

__device__ __forceinline__ int reflect1i(int i, int n) {  // index map of ReflectionPad2d(1): -1 -> 1, n -> n-2
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

// one thread = one float4 (4 channels) of one padded output pixel
__global__ void upcat_pad_fwd_kernel(const float4* __restrict__ a, const float4* __restrict__ skip, float4* __restrict__ y, int B,
                                     int Ca4, int Cs4, int H, int W, int up) {
    const int C4 = Ca4 + Cs4, Hp = H + 2, Wp = W + 2;
    const long long total = (long long)B * Hp * Wp * C4;
    const int Ha = up ? H / 2 : H, Wa = up ? W / 2 : W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int px = (int)(r % Wp);
        r /= Wp;
        const int py = (int)(r % Hp), b = (int)(r / Hp);
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
    const int C4 = Ca4 + Cs4, Hp = H + 2, Wp = W + 2;
    const int Ha = up ? H / 2 : H, Wa = up ? W / 2 : W;
    const long long total = (long long)B * Ha * Wa * Ca4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Ca4);
        long long r = i / Ca4;
        const int ax = (int)(r % Wa);
        r /= Wa;
        const int ay = (int)(r % Ha), b = (int)(r / Ha);
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
    const int C4 = Ca4 + Cs4, Hp = H + 2, Wp = W + 2;
    const long long total = (long long)B * H * W * Cs4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cs4);
        long long r = i / Cs4;
        const int sx = (int)(r % W);
        r /= W;
        const int sy = (int)(r % H), b = (int)(r / H);
        gs[i] = pad_adjoint(gy, (long long)b * Hp * Wp, C4, Ca4 + c, H, W, sy, sx);
    }
}

int grid_for(long long total) {
    long long b = (total + 255) / 256;
    return (int)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

}  // namespace

cudaError_t upcat_pad_fwd(const float* a, const float* skip, float* y, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st) {
    const long long total = (long long)B * (H + 2) * (W + 2) * ((Ca + Cs) / 4);
    upcat_pad_fwd_kernel<<<grid_for(total), 256, 0, st>>>((const float4*)a, (const float4*)skip, (float4*)y, B, Ca / 4, Cs / 4, H, W, up);
    return cudaGetLastError();
}

cudaError_t upcat_pad_bwd(const float* gy, float* ga, float* gskip, int B, int Ca, int Cs, int H, int W, int up, cudaStream_t st) {
    if (ga) {
        const long long total = (long long)B * (up ? H / 2 : H) * (up ? W / 2 : W) * (Ca / 4);
        upcat_pad_bwd_a_kernel<<<grid_for(total), 256, 0, st>>>((const float4*)gy, (float4*)ga, B, Ca / 4, Cs / 4, H, W, up);
    }
    if (gskip && Cs > 0) {
        const long long total = (long long)B * H * W * (Cs / 4);
        upcat_pad_bwd_skip_kernel<<<grid_for(total), 256, 0, st>>>((const float4*)gy, (float4*)gskip, B, Ca / 4, Cs / 4, H, W);
    }
    return cudaGetLastError();
}

}  // namespace mvf
