// Symmetric-memory communication runtime entry points (csrc/comm/symm_comm.cu, csrc/fused/*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace tb {

cudaError_t symm_alloc(size_t bytes, void** ptr);
cudaError_t symm_free(void* ptr);
cudaError_t symm_get_handle(void* ptr, void* handle64);
cudaError_t symm_open_handle(const void* handle64, void** ptr);
cudaError_t symm_close_handle(void* ptr);

// peer_ptrs / pad_ptrs: host arrays of `world` device addresses (this rank's own buffer at index `rank`).
cudaError_t symm_all_gather(const uint64_t* peer_ptrs, const uint64_t* pad_ptrs, size_t src_off_bytes, void* out,
                            size_t bytes, int rank, int world, int channel, uint32_t epoch, uint32_t* block_counter,
                            int num_sms, cudaStream_t stream);
cudaError_t symm_reduce_scatter(const uint64_t* peer_ptrs, const uint64_t* pad_ptrs, size_t src_off_bytes, void* out,
                                size_t n, bool in_bf16, bool out_fp32, float scale, int rank, int world, int channel,
                                uint32_t epoch, uint32_t* block_counter, int num_sms, cudaStream_t stream);
cudaError_t symm_all_to_all(const uint64_t* peer_ptrs, const uint64_t* pad_ptrs, size_t src_off_bytes, void* out,
                            size_t chunk_bytes, int rank, int world, int channel, uint32_t epoch,
                            uint32_t* block_counter, int num_sms, cudaStream_t stream);

}  // namespace tb
