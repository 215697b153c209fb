// Counter-based RNG (Philox4x32-10) for draws generated on device when the
// caller does not supply them (tf.random.uniform / tf.random.normal /
// tf.random.categorical call sites: ibc_agent.py:350, mcmc.py:51,146,249).
// Keyed by (seed, stream id, element index) so results do not depend on the
// launch geometry.
#pragma once

#include <stdint.h>

namespace ibc {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ uint4 philox_at(uint64_t seed, uint64_t stream, uint64_t index) {
  const uint4 ctr = make_uint4((uint32_t)index, (uint32_t)(index >> 32), (uint32_t)stream,
                               (uint32_t)(stream >> 32));
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return philox4x32_10(ctr, key);
}

// [0, 1) with 24 random bits
__device__ __forceinline__ float philox_uniform_float(uint64_t seed, uint64_t stream,
                                                      uint64_t index) {
  const uint4 r = philox_at(seed, stream, index);
  return (float)(r.x >> 8) * (1.0f / 16777216.0f);
}

// [0, 1) with 53 random bits
__device__ __forceinline__ double philox_uniform_double(uint64_t seed, uint64_t stream,
                                                        uint64_t index) {
  const uint4 r = philox_at(seed, stream, index);
  const uint64_t hi = (uint64_t)(r.x >> 5), lo = (uint64_t)(r.y >> 6);
  return (double)((hi << 26) | lo) * (1.0 / 9007199254740992.0);
}

// standard normal (Box-Muller on two 32-bit words)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t stream, uint64_t index) {
  const uint4 r = philox_at(seed, stream, index);
  const float u1 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
  const float u2 = (float)(r.y >> 8) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

}  // namespace ibc
