// GPU input pipeline (datasets/mono_dataset.py:102-184, 206-238): what MonoDataset.preprocess does per training item
// AFTER the resize to the network resolution -- ToTensor, the horizontal flip (mono_dataset.py:224-226), and ColorJitter with one
// parameter set per item applied to all of its frames (mono_dataset.py:228-233: brightness / contrast / saturation in [0.8, 1.2],
// hue in [-0.1, 0.1], in the random order torchvision draws) -- from the item's uint8 frames, so that a step's host-to-device
// copy is the 8-bit frames (1.1 MB per sample at 192x640) instead of six fp32 tensors (8.8 MB).
// Arithmetic: torchvision's tensor kernels (transforms/_functional_tensor.py: _blend, rgb_to_grayscale, _rgb2hsv, _hsv2rgb) in fp32;
// the reference applies the PIL variants, which round to 8 bits after every operation (~1/255 each) and convert to HSV in integers
// (torchvision's own tensor and PIL backends differ by up to 9.6/255 in adjust_hue); parity is stated against the tensor backend.
//   frames  uint8 [B, F, H, W, 3]  (HWC, what PIL / numpy hand over)      color, color_aug  fp32 [F][B, 3, H, W]
//   prm_f   fp32 [B, 4] = brightness, contrast, saturation, hue factors    prm_i  int32 [B, 6] = order[4] (0 b, 1 c, 2 s, 3 h), do_aug, do_flip
// Contrast needs the mean grey level of the WHOLE frame as it is when the operation runs; kernel 1 evaluates the chain up to that
// point and leaves per-CTA partial sums, kernel 2 adds them in a fixed order and runs the full chain.
#include "input.cuh"

namespace mvf {
namespace {

constexpr int NT = 256, NPART = 32;

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float grey(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }
__device__ __forceinline__ void blend(float& r, float& g, float& b, float o, float f) {
    r = clamp01(f * r + (1.f - f) * o);
    g = clamp01(f * g + (1.f - f) * o);
    b = clamp01(f * b + (1.f - f) * o);
}
__device__ __forceinline__ void hue_shift(float& r, float& g, float& b, float f) {
    const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
    const bool eq = maxc == minc;
    const float cr = maxc - minc;
    const float s = cr / (eq ? 1.f : maxc);
    const float div = eq ? 1.f : cr;
    const float rc = (maxc - r) / div, gc = (maxc - g) / div, bc = (maxc - b) / div;
    const float hr = (maxc == r) ? (bc - gc) : 0.f;
    const float hg = ((maxc == g) && (maxc != r)) ? (2.f + rc - bc) : 0.f;
    const float hb = ((maxc != g) && (maxc != r)) ? (4.f + gc - rc) : 0.f;
    float h = fmodf((hr + hg + hb) / 6.f + 1.f, 1.f);
    h = fmodf(h + f, 1.f);
    if (h < 0.f) h += 1.f;   // python's % on a negative operand
    const float v = maxc;
    const float h6 = h * 6.f;
    const float fl = floorf(h6);
    const float fr = h6 - fl;
    const int i = ((int)fl) % 6;
    const float p = clamp01(v * (1.f - s)), q = clamp01(v * (1.f - fr * s)), t = clamp01(v * (1.f - (1.f - fr) * s));
    switch (i) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}
// applies operations order[0 .. n_ops) ; a contrast step uses `mean`
__device__ __forceinline__ void chain(float& r, float& g, float& b, const int* order, const float* f, int n_ops, float mean) {
    for (int k = 0; k < n_ops; ++k) {
        const int op = order[k];
        if (op == 0) blend(r, g, b, 0.f, f[0]);
        else if (op == 1) blend(r, g, b, mean, f[1]);
        else if (op == 2) blend(r, g, b, grey(r, g, b), f[2]);
        else hue_shift(r, g, b, f[3]);
    }
}

// grid (NPART, B * F): partial[bf][part] = sum of the grey level after the operations that precede contrast
__global__ void __launch_bounds__(NT) jitter_grey_partial_kernel(const unsigned char* __restrict__ frames, const float* __restrict__ prm_f,
                                                                 const int* __restrict__ prm_i, float* __restrict__ partial, int F, int H,
                                                                 int W) {
    __shared__ float red[NT / 32];
    const int bf = blockIdx.y, b = bf / F;
    int order[4];
    float f[4];
    for (int k = 0; k < 4; ++k) {
        order[k] = prm_i[6 * b + k];
        f[k] = prm_f[4 * b + k];
    }
    int n_before = 0;
    while (n_before < 4 && order[n_before] != 1) ++n_before;
    const long long n = (long long)H * W, per = (n + NPART - 1) / NPART;
    const long long i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    const unsigned char* src = frames + (size_t)bf * n * 3;
    float s = 0.f;
    for (long long i = i0 + threadIdx.x; i < i1; i += NT) {
        float r = __fdiv_rn((float)src[3 * i], 255.f), g = __fdiv_rn((float)src[3 * i + 1], 255.f), bl = __fdiv_rn((float)src[3 * i + 2], 255.f);
        chain(r, g, bl, order, f, n_before, 0.f);
        s += grey(r, g, bl);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < NT / 32; ++w) t += red[w];
        partial[bf * NPART + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(NT) input_pipeline_kernel(const unsigned char* __restrict__ frames, const float* __restrict__ prm_f,
                                                            const int* __restrict__ prm_i, const float* __restrict__ partial,
                                                            float* const* __restrict__ color, float* const* __restrict__ color_aug, int B,
                                                            int F, int H, int W) {
    const long long n = (long long)H * W, total = (long long)B * F * n;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const long long p = i % n;
        const int bf = (int)(i / n), b = bf / F, fr = bf - b * F;
        const int x = (int)(p % W), y = (int)(p / W);
        const int* pi = prm_i + 6 * b;
        const int sx = pi[5] ? W - 1 - x : x;
        const unsigned char* src = frames + ((size_t)bf * n + (size_t)y * W + sx) * 3;
        float r = __fdiv_rn((float)src[0], 255.f), g = __fdiv_rn((float)src[1], 255.f), bl = __fdiv_rn((float)src[2], 255.f);   // ToTensor: byte / 255, correctly rounded
        float* c = color[fr] + (size_t)b * 3 * n + p;
        c[0] = r; c[n] = g; c[2 * n] = bl;
        if (pi[4]) {
            float mean = 0.f;
            for (int q = 0; q < NPART; ++q) mean += partial[bf * NPART + q];   // fixed order
            mean /= (float)n;
            int order[4] = {pi[0], pi[1], pi[2], pi[3]};
            const float f[4] = {prm_f[4 * b], prm_f[4 * b + 1], prm_f[4 * b + 2], prm_f[4 * b + 3]};
            chain(r, g, bl, order, f, 4, mean);
        }
        float* a = color_aug[fr] + (size_t)b * 3 * n + p;
        a[0] = r; a[n] = g; a[2 * n] = bl;
    }
}

}  // namespace

size_t input_pipeline_workspace_floats(int B, int F) { return (size_t)B * F * NPART; }

cudaError_t input_pipeline(const unsigned char* frames, const float* prm_f, const int* prm_i, float* workspace, float* const* color_dev,
                           float* const* color_aug_dev, int B, int F, int H, int W, cudaStream_t st) {
    jitter_grey_partial_kernel<<<dim3(NPART, B * F), NT, 0, st>>>(frames, prm_f, prm_i, workspace, F, H, W);
    const long long total = (long long)B * F * H * W;
    long long g = (total + NT - 1) / NT;
    if (g > 148 * 16) g = 148 * 16;
    input_pipeline_kernel<<<(unsigned)g, NT, 0, st>>>(frames, prm_f, prm_i, workspace, color_dev, color_aug_dev, B, F, H, W);
    return cudaGetLastError();
}

}  // namespace mvf
