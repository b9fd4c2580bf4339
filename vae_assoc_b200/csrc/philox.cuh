// Philox4x32-10 (Random123, Salmon et al. SC'11) -- device side; oracle twin: oracle/philox.py.
// key = (seed, tag), counter = (row_lo, row_hi, block, step); replaces tf.random_normal (vae_assoc.py:90).
#pragma once
#include <stdint.h>

namespace vaeassoc {

enum PhiloxTag : uint32_t { TAG_EPS = 1, TAG_CODE = 2, TAG_IMG = 3, TAG_JNT = 4, TAG_PRIOR = 5, TAG_PROJ = 16 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = 0xD2511F53ull * (uint64_t)c0;
    const uint64_t p1 = 0xCD9E8D57ull * (uint64_t)c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uint32 -> (0,1), exactly representable: ((x >> 8) + 0.5) * 2^-24
__host__ __device__ __forceinline__ float philox_u01(uint32_t x) {
  return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

#ifdef __CUDACC__
// two words -> two standard normals (Box-Muller; accurate logf / sincospif, no fast-math)
__device__ __forceinline__ void philox_box_muller(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  const float u1 = philox_u01(xa), u2 = philox_u01(xb);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__device__ __forceinline__ void philox_normal4(uint64_t row, uint32_t block, uint32_t step, uint32_t seed,
                                               uint32_t tag, float n[4]) {
  uint32_t w[4];
  philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), block, step, seed, tag, w);
  philox_box_muller(w[0], w[1], n[0], n[1]);
  philox_box_muller(w[2], w[3], n[2], n[3]);
}
#endif

}  // namespace vaeassoc
