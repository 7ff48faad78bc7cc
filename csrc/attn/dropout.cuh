// Counter-based dropout mask shared by the flash-attention forward and backward kernels.
// keep(b*Hq + h, q, k) is a pure function of (seed, head, query row, key) -- both inside their sequence -- so the
// forward (thread = query row, keys along the registers) and the backward (thread = key, queries along the registers)
// regenerate the same mask without storing it.  32-bit integer mix (two multiply-xorshift rounds over the combined
// coordinates); the reference's flash-attn uses Philox4x32 keyed the same way (torchacc/ops/flash_attn.py:313-355 saves
// its rng_state for the backward).  torchacc_b200/ops/attention.py::dropout_keep_mask is the bit-exact PyTorch mirror.
#pragma once
#include <cstdint>

namespace tb {

struct DropoutParams {
  uint32_t thresh24;   // keep iff (hash >> 8) >= thresh24, thresh24 = round(p_drop * 2^24)
  float rp;            // 1 / (1 - p_drop)
  uint32_t seed_lo, seed_hi;
};

__device__ __forceinline__ uint32_t drop_head_part(uint32_t seed_lo, uint32_t bh) { return seed_lo ^ (bh * 0x9E3779B1u); }
__device__ __forceinline__ uint32_t drop_row_part(uint32_t head_part, uint32_t q) {
  return (head_part ^ (q * 0x85EBCA77u)) * 0xC2B2AE3Du;
}
__device__ __forceinline__ uint32_t drop_key_part(uint32_t seed_hi, uint32_t k) { return seed_hi + k * 0x27D4EB2Fu; }
__device__ __forceinline__ bool drop_keep(uint32_t row_part, uint32_t key_part, uint32_t thresh24) {
  uint32_t x = row_part ^ key_part;
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return (x >> 8) >= thresh24;
}

}  // namespace tb
