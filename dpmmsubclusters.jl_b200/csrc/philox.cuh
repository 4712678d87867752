// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), shared by every kernel that needs
// randomness.  Counter = (global point index lo, hi, call counter, stream id), key = 64-bit seed.
// The reference seeds Base.Random identically on every worker (dp-parallel-sampling.jl:37-39) and
// draws rand()/rand(1:2) in point order (utils.jl:29, local_clusters_actions.jl:260,275,478); its
// streams are Julia-version dependent, so they are replaced, not reproduced.  Keying by the GLOBAL
// point index makes every draw independent of how the points are sharded over GPUs.
// The CPU oracle (oracle/dpmm_oracle.py: philox4x32_10 / philox_uniform / philox_bit) restates
// exactly this arithmetic.
#pragma once
#include <cstdint>

#define DPMM_STREAM_LABEL 0u
#define DPMM_STREAM_SUBLABEL 1u
#define DPMM_STREAM_RANDBITS 2u
#define DPMM_STREAM_INIT 3u
#define DPMM_STREAM_GUMBEL 4u

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ Philox4 philox_draw(uint64_t seed, uint32_t stream,
                                                       uint32_t call, uint64_t gidx) {
  return philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), call, stream, (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}

// 53-bit uniform in [0,1): stand-in for Julia's rand()::Float64.
__host__ __device__ __forceinline__ double philox_to_uniform(const Philox4& r) {
  const double hi = (double)(r.x >> 5);  // 27 bits
  const double lo = (double)(r.y >> 6);  // 26 bits
  return (hi * 67108864.0 + lo) * (1.0 / 9007199254740992.0);
}
