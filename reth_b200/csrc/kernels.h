// kernels.h — host-side launch interface of the CUDA kernels (internal to libb200trie.so).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

cudaError_t launch_keccak256_fixed(const void *d_in, uint32_t msg_len, uint32_t stride, uint64_t n, void *d_out,
                                   cudaStream_t s, unsigned *launches);
cudaError_t launch_keccak256_var(const void *d_data, const void *d_offsets, uint64_t n, void *d_out, cudaStream_t s,
                                 unsigned *launches);

}  // namespace b200
