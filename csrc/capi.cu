// Flat C ABI over the native kernels, loaded from Python with ctypes (torchacc_b200/_native.py).
// Every function returns a cudaError_t as int; pointers are raw device addresses (tensor.data_ptr()) and
// `stream` is a cudaStream_t handle (torch.cuda.current_stream().cuda_stream).
#include <cuda_runtime.h>
#include <stdint.h>

#include "attn/attn.h"
#include "comm/comm.h"
#include "gemm/gemm.h"
#include "ops/ops.h"

#define TB_API extern "C" __attribute__((visibility("default")))

static inline cudaStream_t S(uint64_t s) { return reinterpret_cast<cudaStream_t>(s); }
template <typename T>
static inline T* P(uint64_t p) { return reinterpret_cast<T*>(p); }

TB_API const char* tb_error_string(int code) { return cudaGetErrorString(static_cast<cudaError_t>(code)); }
TB_API int tb_abi_version() { return 3; }

TB_API int tb_device_info(int device, int* num_sms, int* cc_major, int* cc_minor, long long* smem_optin) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return (int)e;
  *num_sms = prop.multiProcessorCount;
  *cc_major = prop.major;
  *cc_minor = prop.minor;
  *smem_optin = (long long)prop.sharedMemPerBlockOptin;
  return 0;
}

// ---- GEMM -------------------------------------------------------------------------------------------
TB_API int tb_gemm_bf16(uint64_t A, uint64_t B, uint64_t D, uint64_t bias, int M, int N, int K, long long lda,
                        long long ldb, long long ldd, int a_mn_major, int b_mn_major, int out_fp32, int accumulate,
                        int cluster, int num_sms, uint64_t stream) {
  return (int)tb::gemm_bf16(P<void>(A), P<void>(B), P<void>(D), P<void>(bias), M, N, K, lda, ldb, ldd,
                            a_mn_major != 0, b_mn_major != 0, out_fp32 != 0, accumulate != 0, cluster, num_sms,
                            S(stream));
}

// ---- norm / rope / activation -----------------------------------------------------------------------
TB_API int tb_rmsnorm_fwd(uint64_t x, uint64_t res, uint64_t w, uint64_t y, uint64_t h_out, uint64_t rstd, int rows,
                          int H, float eps, int num_sms, uint64_t stream) {
  return (int)tb::rmsnorm_fwd(P<void>(x), P<void>(res), P<void>(w), P<void>(y), P<void>(h_out), P<float>(rstd), rows,
                              H, eps, num_sms, S(stream));
}
TB_API int tb_rmsnorm_bwd(uint64_t dy, uint64_t x, uint64_t w, uint64_t rstd, uint64_t dres, uint64_t dx, uint64_t dw,
                          int rows, int H, int num_sms, uint64_t stream) {
  return (int)tb::rmsnorm_bwd(P<void>(dy), P<void>(x), P<void>(w), P<float>(rstd), P<void>(dres), P<void>(dx),
                              P<float>(dw), rows, H, num_sms, S(stream));
}
TB_API int tb_rope_inplace(uint64_t x, uint64_t cos_t, uint64_t sin_t, uint64_t positions, long long T, int nheads,
                           int D, long long token_stride, int seq_len, int backward, int num_sms, uint64_t stream) {
  return (int)tb::rope_inplace(P<void>(x), P<float>(cos_t), P<float>(sin_t), P<int>(positions), T, nheads, D,
                               token_stride, seq_len, backward != 0, num_sms, S(stream));
}
TB_API int tb_swiglu_fwd(uint64_t g, uint64_t u, uint64_t h, long long T, int F, long long ldg, long long ldu,
                         int num_sms, uint64_t stream) {
  return (int)tb::swiglu_fwd(P<void>(g), P<void>(u), P<void>(h), T, F, ldg, ldu, num_sms, S(stream));
}
TB_API int tb_swiglu_bwd(uint64_t dh, uint64_t g, uint64_t u, uint64_t dg, uint64_t du, long long T, int F,
                         long long ldg, long long ldu, long long lddg, long long lddu, int num_sms, uint64_t stream) {
  return (int)tb::swiglu_bwd(P<void>(dh), P<void>(g), P<void>(u), P<void>(dg), P<void>(du), T, F, ldg, ldu, lddg,
                             lddu, num_sms, S(stream));
}

// ---- loss / optimizer -------------------------------------------------------------------------------
TB_API int tb_cross_entropy(uint64_t logits, uint64_t labels, uint64_t loss_rows, uint64_t lse_rows, int n, int V,
                            long long ld, int ignore_index, uint64_t scale_ptr, float scale_val, int write_grad,
                            uint64_t stream) {
  return (int)tb::cross_entropy_fwd_bwd(P<void>(logits), P<long long>(labels), P<float>(loss_rows), P<float>(lse_rows),
                                        n, V, ld, ignore_index, P<float>(scale_ptr), scale_val, write_grad != 0,
                                        S(stream));
}
TB_API int tb_adamw_flat(uint64_t p, uint64_t g, int grad_is_bf16, uint64_t m, uint64_t v, uint64_t p_lp, long long n,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                         uint64_t grad_scale, uint64_t found_inf, int num_sms, uint64_t stream) {
  return (int)tb::adamw_flat(P<float>(p), P<void>(g), grad_is_bf16 != 0, P<float>(m), P<float>(v), P<void>(p_lp), n,
                             lr, beta1, beta2, eps, weight_decay, step, P<float>(grad_scale), P<float>(found_inf),
                             num_sms, S(stream));
}
TB_API int tb_sqnorm_accumulate(uint64_t g, int is_bf16, long long n, uint64_t out, float pre_scale, int num_sms,
                                uint64_t stream) {
  return (int)tb::sqnorm_accumulate(P<void>(g), is_bf16 != 0, n, P<float>(out), pre_scale, num_sms, S(stream));
}
TB_API int tb_scale_inplace(uint64_t g, int is_bf16, long long n, uint64_t scale, int num_sms, uint64_t stream) {
  return (int)tb::scale_inplace(P<void>(g), is_bf16 != 0, n, P<float>(scale), num_sms, S(stream));
}
