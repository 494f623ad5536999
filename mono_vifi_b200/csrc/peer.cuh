// Cross-GPU sum of short float64 vectors over NVLink peer memory (see peer.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mvf {
constexpr int PEER_SLOTS = 4, PEER_MAX_WORLD = 16, PEER_MAX_N = 2056, PEER_CHANNELS = 8;
constexpr size_t PEER_CHANNEL_BYTES = (size_t)PEER_SLOTS * PEER_MAX_WORLD * 8 + (size_t)PEER_SLOTS * PEER_MAX_WORLD * PEER_MAX_N * 8;
size_t peer_buffer_bytes();
cudaError_t peer_allreduce_f64(double* vec, int n, void* const* peers_dev, int rank, int world, int channel, unsigned long long* seq_local,
                               cudaStream_t st);
}  // namespace mvf
