// slam_philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11), keyed exactly as
// oracle.philox_uniform: counter = (sample_lo, sample_hi, j/2, TAG), key = (seed_lo, seed_hi);
// words (w0,w1) give parameter 2*(j/2), (w2,w3) give 2*(j/2)+1; u = ((hi<<32 | lo) >> 11) * 2^-53.
// Replaces the legacy global np.random streams of basis.py:111 / basisv2.py:157-167 with a stream
// that any rank can regenerate for its own shard.
#pragma once
#include <stdint.h>

namespace slam {

constexpr uint32_t kPhiloxTag = 0x51A3B200u;

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

// uniform [0,1) for (seed, sample, parameter j)
__device__ __forceinline__ double philox_u53(uint64_t seed, uint64_t sample, int j) {
  uint32_t c[4] = {(uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)(j >> 1), kPhiloxTag};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (j & 1) ? u53(c[2], c[3]) : u53(c[0], c[1]);
}

// both uniforms of the Philox block that serves parameters 2*jj and 2*jj + 1
__device__ __forceinline__ void philox_u53_pair(uint64_t seed, uint64_t sample, int jj, double* u0, double* u1) {
  uint32_t c[4] = {(uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)jj, kPhiloxTag};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  *u0 = u53(c[0], c[1]);
  *u1 = u53(c[2], c[3]);
}

// lo + span*u with a separate multiply and add (numpy does not fuse; keeps the stream bit-identical)
__device__ __forceinline__ double philox_param(uint64_t seed, uint64_t sample, int j, double lo, double span) {
  return __dadd_rn(lo, __dmul_rn(span, philox_u53(seed, sample, j)));
}

}  // namespace slam
