// Public (C++) entry points of the tcgen05 GEMM family.  Raw pointers + stream so the kernels build
// without any PyTorch headers; csrc/bindings.cpp wraps them as torch ops.
#pragma once
#include <cuda_runtime.h>

namespace tb {

// D[M,N] (+)= A_op[M,K] * B_op[N,K]^T (+ bias).  bf16 inputs, fp32 accumulation in TMEM.
//   a_mn_major == false : A stored [M][K] (row pitch lda);  true : A stored [K][M]
//   b_mn_major == false : B stored [N][K] (row pitch ldb);  true : B stored [K][N]
//   out_fp32            : D is float32 instead of bf16
//   accumulate          : D += result (gradient accumulation)
//   cluster             : 1 = one CTA per 128x256 tile, 2 = CTA pair per 256x256 tile (cta_group::2)
cudaError_t gemm_bf16(const void* A, const void* B, void* D, const void* bias, int M, int N, int K, long long lda,
                      long long ldb, long long ldd, bool a_mn_major, bool b_mn_major, bool out_fp32, bool accumulate,
                      int cluster, int num_sms, cudaStream_t stream);

}  // namespace tb
