// Flash-attention entry points (csrc/attn/flash_fwd.cu, flash_bwd.cu).
// Tensors are token-major [tokens, heads, D] bf16 or fp16 (is_bf16) with an arbitrary token stride (so q/k/v may be column blocks of
// one fused QKV activation); sequences are either fixed length (B x S) or packed with int32 cu_seqlens [B+1].
// alibi_slopes: optional fp32 [Hq] (alibi_batch_stride == 0) or [B, Hq] (stride Hq): bias -slope * |i + Sk - Sq - j|.
#pragma once
#include <cuda_runtime.h>

namespace tb {

cudaError_t flash_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const int* cu_q,
                           const int* cu_k, int B, int Sq, int Sk, int Hq, int Hk, int D, long long q_ts,
                           long long k_ts, long long v_ts, long long o_ts, float scale, bool causal, int wl, int wr,
                           long long Tq, long long Tk, int max_q_len, bool is_bf16, const float* alibi_slopes,
                           int alibi_batch_stride, cudaStream_t stream);

cudaError_t flash_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                           const float* lse, void* dq, void* dk, void* dv, float* dq_acc, float* delta,
                           const int* cu_q, const int* cu_k, int B, int Sq, int Sk, int Hq, int Hk, int D,
                           long long q_ts, long long k_ts, long long v_ts, long long do_ts, float scale, bool causal,
                           int wl, int wr, long long Tq, long long Tk, long long dq_ts, long long dk_ts,
                           long long dv_ts, int num_sms, bool is_bf16, const float* alibi_slopes,
                           int alibi_batch_stride, cudaStream_t stream);

// Debug: when non-null, CTA (0,0,0) of the backward kernel writes clock64() stamps [64 iterations][16 slots].
void flash_attn_bwd_set_trace(long long* p);
void flash_attn_fwd_set_trace(long long* p);

}  // namespace tb
