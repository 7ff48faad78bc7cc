// Public (C++) entry points of the tcgen05 GEMM family.  Raw pointers + stream so the kernels build
// without any PyTorch headers; csrc/bindings.cpp wraps them as torch ops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

// D[M,N] (+)= A_op[M,K] * B_op[N,K]^T (+ bias).  bf16 (or, with is_fp16, IEEE fp16) inputs, fp32 accumulation in TMEM.
//   a_mn_major == false : A stored [M][K] (row pitch lda);  true : A stored [K][M]
//   b_mn_major == false : B stored [N][K] (row pitch ldb);  true : B stored [K][N]
//   out_fp32            : D is float32 instead of bf16
//   accumulate          : D += result (gradient accumulation)
//   cluster             : 1 = one CTA per 128x256 tile, 2 = CTA pair per 256x256 tile (cta_group::2)
cudaError_t gemm_bf16(const void* A, const void* B, void* D, const void* bias, int M, int N, int K, long long lda,
                      long long ldb, long long ldd, bool a_mn_major, bool b_mn_major, bool out_fp32, bool accumulate,
                      int cluster, int num_sms, cudaStream_t stream, bool is_fp16 = false);

// Tile scheduling of the unfused GEMM: 1 = dynamic (one cluster per tile, claimed with cluster launch control; robust
// against SMs that are busy with a concurrent kernel), 0 = static striping over a persistent grid.  mode < 0 queries.
int gemm_sched_mode(int mode);

// Same with a separate addend: D = A_op * B_op^T (+ bias) + C, C with D's dtype and row pitch ldc (C may alias D).
// Used to fuse the transformer residual add into the down-projection epilogue.
cudaError_t gemm_bf16_ex(const void* A, const void* B, void* D, const void* bias, const void* C, int M, int N, int K,
                         long long lda, long long ldb, long long ldd, long long ldc, bool a_mn_major, bool b_mn_major,
                         bool out_fp32, int cluster, int num_sms, cudaStream_t stream, bool is_fp16 = false);

// gate|up projection with the SwiGLU activation in the epilogue: H[M][F] = silu(A Wg^T) * (A Wu^T), W_gu = [Wg; Wu]
// stored [2F][K]; GU (optional, [M][2F]) receives the pre-activation for the backward pass.  F % 128 == 0.
cudaError_t gemm_swiglu_bf16(const void* A, const void* Wgu, void* H, void* GU, int M, int F, int K, long long lda,
                             long long ldb, long long ldh, long long ldgu, int num_sms, cudaStream_t stream,
                             bool is_fp16 = false);

// ---- fused tensor-parallel kernels (see the FuseArgs comment in gemm_bf16.cu) ----
// all-gather -> GEMM: D[world*rows, N] = gather(A)[world*rows, K] * B_op^T.  `a_full` is this rank's symmetric
// gathered buffer whose own row block is already filled; `peer_a_full[r]` is rank r's mapping of the same buffer.
cudaError_t ag_gemm_bf16(const uint64_t* peer_a_full, const uint64_t* pad_ptrs, void* a_full, const void* B, void* D,
                         const void* bias, int rows_per_rank, int N, int K, long long ldb, long long ldd,
                         bool b_mn_major, int rank, int world, uint32_t* flags, uint32_t flag_target,
                         uint32_t* block_counter, int channel, uint32_t epoch, int comm_clusters, int num_sms,
                         cudaStream_t stream);
// GEMM -> reduce-scatter, part 1: partial tiles are stored into the owner rank's staging slot [rank] (peer stores)
// and counted in peer_counters[owner][rank] (4 arrivals per 128-row CTA tile).
cudaError_t gemm_rs_bf16(const void* A, const void* B, const uint64_t* peer_stage, const uint64_t* peer_counters,
                         const uint64_t* pad_ptrs, int rows_per_rank, int N, int K, long long lda, long long ldb,
                         bool a_mn_major, bool b_mn_major, int rank, int world, int channel, uint32_t epoch,
                         int num_sms, cudaStream_t stream);
// part 2: out = sum over sources of stage[src] (+ residual) once counters[src] >= expected for every source.
cudaError_t rs_reduce_bf16(const void* stage, const uint32_t* counters, uint32_t expected, const void* residual,
                           void* out, int rows, int N, int world, long long slot_stride, int num_sms, cudaStream_t stream);

// Block-scaled FP8 GEMM (gemm_mxfp8.cu): D[M,N] = (A[M,K] * SFA) (B[N,K] * SFB)^T (+ C); e4m3 operands, UE8M0 scales per
// 32 elements of K in the atom-tiled layout written by ops/quant_mxfp8.cu; K % 128 == 0.
cudaError_t gemm_mxfp8(const void* A, const void* sfa, const void* B, const void* sfb, void* D, const void* C, int M,
                       int N, int K, long long lda, long long ldb, long long ldd, long long ldc, bool out_fp32,
                       int num_sms, cudaStream_t stream);

}  // namespace tb
