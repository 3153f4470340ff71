// Counter-based dropout stream: Philox4x32-10 (Salmon et al., SC'11 -- public algorithm).
// keep(site, idx) is a pure function of (seed, offset, site, element index), so the backward
// regenerates the forward's mask without storing it.
#pragma once
#include <stdint.h>

namespace iisan {

struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// counter = (idx_lo, idx_hi, site, offset_lo) ; key = (seed_lo, seed_hi ^ offset_hi)
__host__ __device__ inline void philox4x32_10(uint64_t seed, uint64_t offset, uint32_t site, uint64_t idx4, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)idx4, (uint32_t)(idx4 >> 32), site, (uint32_t)offset};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// keep-probability test for element `idx` of dropout site `site`: u = r * 2^-32 ; keep iff u >= p, i.e. r >= thr(p)
__host__ __device__ inline uint32_t dropout_threshold(float p) {
  return (uint32_t)((double)p * 4294967296.0 > 4294967295.0 ? 4294967295.0 : (double)p * 4294967296.0);
}
__host__ __device__ inline bool dropout_keep_thr(uint64_t seed, uint64_t offset, uint32_t site, uint64_t idx, uint32_t thr) {
  uint32_t r[4];
  philox4x32_10(seed, offset, site, idx >> 2, r);
  return r[idx & 3] >= thr;
}
__host__ __device__ inline bool dropout_keep(uint64_t seed, uint64_t offset, uint32_t site, uint64_t idx, float p) {
  return dropout_keep_thr(seed, offset, site, idx, dropout_threshold(p));
}

}  // namespace iisan
