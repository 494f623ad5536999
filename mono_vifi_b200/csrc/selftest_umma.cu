// Device self-test of the tcgen05 plumbing used by the convolution kernels: one CTA computes
//   D[128, N] = A . B^T   with B given as [N][K] (K-major) and A given either as [128][K] (K-major) or as [K][128]
//   (MN-major, the layout NCHW activations have), TF32 inputs, fp32 accumulate.
// tests/test_conv_tc_cuda.py compares it with a plain matmul; it pins the descriptor encodings in tc_common.cuh.
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace mvf {
namespace tc {
namespace {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB, float* D, int N, int K,
                                                            int a_mn_major, int row_off, int base_off_mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* smA = smem;           // 20 KB: 160 x 32 floats (K-major) / 16 KB + slack (MN-major)
    unsigned char* smB = smem + 20480;   // up to 32 KB: N x 32 floats
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 20480 + 32768);
    uint64_t* mma_bar = full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(full_bar, 1);
        mbar_init(mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, N, a_mn_major, 0);
    const int nkb = K / 32;
    for (int kb = 0; kb < nkb; ++kb) {
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(full_bar, 20480 + N * 128);
            if (a_mn_major) {
                // A given as [K+8][128]: dims (m%32, k, m/32), box (32, 40, 4) -> smem [m/32][40 k-rows][m%32]
                tma_load_3d(smA, &mapA, full_bar, 0, kb * 32, 0);
            } else {
                tma_load_3d(smA, &mapA, full_bar, 0, 0, kb);  // dims (k%32, m, k/32), box (32, 160, 1)
            }
            tma_load_3d(smB, &mapB, full_bar, 0, 0, kb);
            mbar_wait(full_bar, kb & 1);
            tc_fence_after();
            for (int kg = 0; kg < 4; ++kg) {
                uint64_t adesc, bdesc;
                if (a_mn_major)  // k-rows [row_off + 8 kg, +8) of each 32-wide MN atom; atoms 40 rows (5120 B) apart
                    adesc = make_smem_desc(smem_u32(smA) + (row_off + 8 * kg) * 128, 40 * 128, 512, SWZ_128B_BASE32B);
                else {
                    // experiment: operand = rows [row_off, row_off + 128) of the 160 loaded rows (start not 1024-aligned)
                    const uint32_t start = smem_u32(smA) + row_off * 128 + kg * 32;
                    adesc = make_smem_desc(start, 16, 1024, SWZ_128B);
                    if (base_off_mode == 1) adesc |= (uint64_t)((start >> 7) & 7) << 49;
                }
                bdesc = make_smem_desc(smem_u32(smB) + kg * 32, 16, 1024, SWZ_128B);
                umma_tf32(tmem_d, adesc, bdesc, idesc, (kb > 0 || kg > 0) ? 1u : 0u);
            }
            umma_commit(mma_bar);
            mbar_wait(mma_bar, kb & 1);  // single-buffered: wait until the MMAs have consumed the tiles
        }
        __syncthreads();
    }
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
}  // namespace

// A: [160][K] of which rows [row_off, row_off+128) are used (a_mn_major = 0) or [K+8][128] of which k-rows
// [row_off, row_off+K) are used (a_mn_major = 1); B: [N][K];
// D: [128][N].  K % 32 == 0, N % 16 == 0, N <= 256.
cudaError_t umma_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn_major, cudaStream_t st,
                          int row_off, int base_off_mode) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess || !fp)
        return cudaErrorNotSupported;
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
    CUtensorMap mapA, mapB;
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (a_mn_major) {
        cuuint64_t dims[4] = {32, (cuuint64_t)(K + 8), 4, 1};
        cuuint64_t strides[3] = {128 * 4, 32 * 4, 128 * 4};
        cuuint32_t box[4] = {32, 40, 4, 1};
        if (enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(A), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    } else {
        cuuint64_t dims[3] = {32, 160, (cuuint64_t)(K / 32)};  // A holds 160 rows; the product uses 128 of them
        cuuint64_t strides[2] = {(cuuint64_t)K * 4, 128};
        cuuint32_t box[3] = {32, 160, 1};
        if (enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(A), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    {
        cuuint64_t dims[3] = {32, (cuuint64_t)N, (cuuint64_t)(K / 32)};
        cuuint64_t strides[2] = {(cuuint64_t)K * 4, 128};
        cuuint32_t box[3] = {32, (cuuint32_t)N, 1};
        if (enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(B), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    const int smem = 20480 + 32768 + 1024 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    umma_selftest_kernel<<<1, 128, smem, st>>>(mapA, mapB, D, N, K, a_mn_major, row_off, base_off_mode);
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace mvf
