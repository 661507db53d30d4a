// Counter-hash dropout RNG shared by the mask generator (misc.cu: dasa_dropout_mask / _dev) and by the kernels that draw their
// keep flags in place instead of reading a materialised mask (forward-only dropout sites of the frozen transformer stack:
// mha_h16.cu, dropout_residual_layernorm in encoder.cu). One stream = (seed, base): BYTE e of the stream is 16-bit lane e & 3 of
// hash64(mix_seed(seed), base + (e >> 2)); keep iff that 16-bit uniform >= p * 65536. dasa_dropout_mask(mask, n, p, seed, base)
// with n % 16 == 0 writes exactly these bytes, which is how the tests check the in-place draws against a materialised mask.
#pragma once
#include <stdint.h>

// The seed goes through its own avalanche round before the element index is folded in. (Adding the raw seed to
// (idx + 1) * G made "seed + G" the same stream shifted by one element: the per-replay seed bump of a captured graph produced
// masks correlated with the previous iteration's.)
__device__ __forceinline__ uint64_t mix_seed(uint64_t seed) {
  uint64_t s = (seed ^ 0x2545F4914F6CDD1Dull) * 0xD6E8FEB86659FD93ull;
  s = (s ^ (s >> 32)) * 0xD6E8FEB86659FD93ull;
  return s ^ (s >> 32);
}

// splitmix64-style finaliser over mixed_seed ^ counter
__device__ __forceinline__ uint64_t hash64(uint64_t mixed_seed, uint64_t idx) {
  uint64_t z = mixed_seed ^ ((idx + 1) * 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 65536.0f); }

struct DropStream {
  uint64_t mixed;      // mix_seed(seed)
  uint64_t base;       // hash index of stream byte 0
  uint32_t thr;
};

// keep flags of stream bytes 4w .. 4w+3 in bits 0..3
__device__ __forceinline__ uint32_t stream_keep4(const DropStream& s, uint64_t w) {
  const uint64_t h = hash64(s.mixed, s.base + w);
  return ((uint32_t)(h & 0xFFFF) >= s.thr ? 1u : 0u) | ((uint32_t)((h >> 16) & 0xFFFF) >= s.thr ? 2u : 0u) |
         ((uint32_t)((h >> 32) & 0xFFFF) >= s.thr ? 4u : 0u) | ((uint32_t)(h >> 48) >= s.thr ? 8u : 0u);
}
