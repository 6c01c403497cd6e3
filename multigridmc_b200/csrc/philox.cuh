// Counter-based noise for the Gibbs sweeps: Philox4x32-10 (Salmon et al., SC'11) generated in
// registers, two N(0,1) per call via Box-Muller on two 53-bit uniforms.
//
// Replaces the reference's sequential std::mt19937_64 + std::normal_distribution stream
// (sampler/sampler.hh:69-71, sor_sampler.cc:42-46).  Counter layout (DESIGN.md "Noise"):
//   key  = 64-bit seed
//   c0   = stream-local index:
//            site noise      ((j * G + (i >> 2)) << 1) | (i & 1),  G = nx/4 + 1; the pair of
//                            same-colour sites (i, i+2) of an aligned group of 4 columns shares
//                            one call: normal = (i & 2) ? z1 : z0
//            low-rank noise  0x80000000 | (k >> 1), normal = (k & 1) ? z1 : z0
//            coarse Cholesky 0x40000000 | (ell >> 1), normal = (ell & 1) ? z1 : z0
//   c1   = (level << 24) | sweep counter within the sample
//   c2   = sample index,  c3 = global chain id
// so that a site's noise is a pure function of (seed, chain, sample, level, sweep, site): tiles may
// recompute halo sites redundantly and any domain decomposition reproduces the same chain.
#pragma once
#include <cstdint>

namespace mgmc {

struct PhiloxKey {
  uint32_t k0, k1;
};

__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// two independent N(0,1) variates from one counter
__device__ __forceinline__ void normal_pair(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double &z0, double &z1) {
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t a = (uint64_t)c0 | ((uint64_t)c1 << 32), b = (uint64_t)c2 | ((uint64_t)c3 << 32);
  const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

}  // namespace mgmc
