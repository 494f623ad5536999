// Latency-bound cross-GPU sum of a short float64 vector over NVLink peer memory: the exchange inside SyncBatchNorm
// (train.py:205-208: nn.SyncBatchNorm.convert_sync_batchnorm makes every BatchNorm of every network reduce its batch statistics over
// all ranks -- 240 tiny all-reduces per step of the ResNet18 configuration, each on the critical path of its layer).
// An NCCL all-reduce of 1 KB costs 10-20 us of launch + protocol latency per call; here ONE single-CTA kernel per exchange
//   1. pushes this rank's vector into a slot of EVERY rank's symmetric buffer (remote stores through NVLink / NVSwitch),
//   2. releases a per-(slot, source rank) flag carrying the exchange's sequence number on every peer,
//   3. waits for the flags of all ranks in its OWN buffer, and
//   4. adds the world's contributions in rank order (the same order on every rank: identical bits everywhere).
// No host involvement, no communicator: the kernel is an ordinary graph node, so it replays with the captured step.  Exchanges issued
// from different CUDA streams use different CHANNELS (own slots, flags and sequence counter), so concurrent branches of the step
// neither serialise nor depend on a common issue order across ranks.  A rank can run at most one exchange ahead of the slowest rank
// of a channel (it needs that rank's contribution to finish the next one), so a ring of 4 slots is never overwritten while in use.
#include "peer.cuh"

#include <cstdint>

namespace mvf {
namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// symmetric buffer of one rank: [channel][ flags[PEER_SLOTS][PEER_MAX_WORLD] u64 | data[PEER_SLOTS][PEER_MAX_WORLD][PEER_MAX_N] f64 ]
__device__ __forceinline__ unsigned long long* flags_of(unsigned char* base, int ch) {
    return reinterpret_cast<unsigned long long*>(base + (size_t)ch * PEER_CHANNEL_BYTES);
}
__device__ __forceinline__ double* data_of(unsigned char* base, int ch) {
    return reinterpret_cast<double*>(base + (size_t)ch * PEER_CHANNEL_BYTES + (size_t)PEER_SLOTS * PEER_MAX_WORLD * 8);
}

__global__ void __launch_bounds__(512) peer_allreduce_f64_kernel(double* __restrict__ vec, int n, unsigned char* const* __restrict__ peers,
                                                                 int rank, int world, int ch, unsigned long long* __restrict__ seq_local) {
    const unsigned long long seq = *seq_local + 1;
    const int slot = (int)(seq % PEER_SLOTS);
    // 1. push
    for (int r = 0; r < world; ++r) {
        double* dst = data_of(peers[r], ch) + ((size_t)slot * PEER_MAX_WORLD + rank) * PEER_MAX_N;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vec[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. + 3. one thread per peer: release my flag there, then wait for that peer's flag here
    if (threadIdx.x < world) {
        st_release_sys(flags_of(peers[threadIdx.x], ch) + slot * PEER_MAX_WORLD + rank, seq);
        const unsigned long long* mine = flags_of(peers[rank], ch) + slot * PEER_MAX_WORLD + threadIdx.x;
        while (ld_acquire_sys(mine) < seq) {
        }
    }
    __syncthreads();
    // 4. fixed-order sum
    const double* src = data_of(peers[rank], ch) + (size_t)slot * PEER_MAX_WORLD * PEER_MAX_N;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += ld_volatile_f64(src + (size_t)r * PEER_MAX_N + i);
        vec[i] = s;
    }
    if (threadIdx.x == 0) *seq_local = seq;
}

}  // namespace

size_t peer_buffer_bytes() { return (size_t)PEER_CHANNELS * PEER_CHANNEL_BYTES; }

cudaError_t peer_allreduce_f64(double* vec, int n, void* const* peers_dev, int rank, int world, int channel, unsigned long long* seq_local,
                               cudaStream_t st) {
    peer_allreduce_f64_kernel<<<1, 512, 0, st>>>(vec, n, reinterpret_cast<unsigned char* const*>(peers_dev), rank, world, channel, seq_local);
    return cudaGetLastError();
}

}  // namespace mvf
