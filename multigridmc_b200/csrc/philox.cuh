// Counter-based noise for the Gibbs sweeps: Philox4x32-7 (Salmon et al., SC'11; rounds: see kPhiloxRounds) generated in
// registers, two N(0,1) per call via Box-Muller on two 52-bit uniforms.
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
//
// The transcendental part is hand-written (table-driven kernels, ~1 ulp, see below) because the sweeps
// are otherwise instruction-issue bound: CUDA's log / sincospi / sqrt carry special-case paths that a
// uniform in (0,1) never takes.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define MGMC_HD __host__ __device__ __forceinline__
#else
#define MGMC_HD inline
#endif

namespace mgmc {

// ------------------------------------------------------------------------------------------------
// Normal variates: Box-Muller on two 52-bit uniforms u = (k + 1/2) 2^-52,
//   z0 = sqrt(-2 ln u1) cos(2 pi u2),  z1 = sqrt(-2 ln u1) sin(2 pi u2),
// evaluated to ~1 ulp with two 32-entry tables (512 bytes each; the tile kernel stages them in shared memory)
// instead of the long fdlibm kernels -- the sweeps are instruction-issue bound:
//   * -2 ln u: u = 2^e m; cell j = top 5 mantissa bits, r = m / c_j - 1 (one FMA with the tabulated 1 / c_j,
//     |r| <= 2^-6), -2 ln u = e' (-2 ln 2) + T2[j] - 2 log1p(r) with the degree-9 Taylor polynomial (remainder
//     < 2^-57 relative).  Cells with m > 1.43 use m / 2 and e + 1, and the last cell has c = 2 exactly, so that
//     u -> 1 (e' = 0, T2 = 0) suffers no cancellation.
//   * sin / cos(2 pi u): cell k = top 5 bits, tabulated (sin, cos) at the cell centre, angle addition with
//     |y| <= pi / 32: sin y up to y^9, cos y - 1 up to y^8 (remainders < 2^-55); no quadrant logic.  The offset
//     from the cell centre is formed exactly from the integer bits.
//   * sqrt: MUFU.RSQ64H seed + two coupled Goldschmidt steps.
// Host and device run the same code on the same tables (normal_tables.inc <- tools/gen_normal_tables.py).
// ------------------------------------------------------------------------------------------------
// [2 j], [2 j + 1]: 1 / c_j, T2[j];   [64 + 2 k], [64 + 2 k + 1]: sin, cos at the centre of angle cell k
static const double kNormalTabHost[128] = {
#include "normal_tables.inc"
};
#if defined(__CUDACC__)
static __device__ const double kNormalTabDev[128] = {
#include "normal_tables.inc"
};
#endif

MGMC_HD double bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d;
  std::memcpy(&d, &b, sizeof(d));
  return d;
#endif
}
MGMC_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t b;
  std::memcpy(&b, &d, sizeof(d));
  return b;
#endif
}
MGMC_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
// sqrt(t) for t in [2^-52, 80]: 22-bit reciprocal square root seed, two coupled Goldschmidt steps
// (g -> sqrt t, h -> 1 / (2 sqrt t); the error is squared per step: 2^-22 -> 2^-43 -> rounding)
MGMC_HD double sqrt_(double t) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));  // MUFU.RSQ64H
#else
  const double y = (double)(1.0f / std::sqrt((float)t));
#endif
  double g = t * y, h = 0.5 * y;
  double rho = fma_(-g, h, 0.5);
  g = fma_(g, rho, g);
  h = fma_(h, rho, h);
  rho = fma_(-g, h, 0.5);
  return fma_(g, rho, g);
}

// The constants that need all 64 bits travel in the kernel parameters (one LDCU.64 per use; as literals
// they would cost two UMOVs each).  Every other coefficient below has zero low 32 bits (high-order
// coefficients are rounded to that form, which moves the result by < 2^-60) and is an immediate operand.
struct NormalConsts {
  double v[11];
};
enum { NC_U0 = 0, NC_L3, NC_L5, NC_L6, NC_KLO, NC_ASCALE, NC_S3, NC_S5, NC_S7, NC_C4, NC_C6 };
#define MGMC_NORMAL_CONSTS                                                                                         \
  {{1.0 - 1.1102230246251565e-16, /* 1 - 2^-53 */                                                                  \
    -0.6666666666666666, -0.4, 0.3333333333333333, /* -2/3, -2/5, 1/3 */                                           \
    3.809308599915536e-09,                         /* -2 ln 2 - K_hi */                                            \
    6.975736996017264e-16,                         /* 2 pi / 2^53 */                                               \
    -0.16666666666666666, 0.008333333333333333, -0.0001984126984126984, /* -1/3!, 1/5!, -1/7! */                   \
    0.041666666666666664, -0.001388888888888889}}  /* 1/4!, -1/6! */
static const NormalConsts kNormalConstsHost = MGMC_NORMAL_CONSTS;

// -2 ln(u) for u = (k + 1/2) 2^-52, k = top 52 bits of a
MGMC_HD double neg2log_u52(uint64_t a, const NormalConsts &mc, const double *tab) {
  // (1 + k 2^-52) - (1 - 2^-53): exact, normalised by the adder
  const double u = bits_to_double(0x3FF0000000000000ull | (a >> 12)) - mc.v[NC_U0];
  const uint64_t ub = double_to_bits(u);
  const uint32_t hx = (uint32_t)(ub >> 32);
  const uint32_t j = (hx >> 15) & 31u;
  const int e = (int)(hx >> 20) - 1023 + (int)((j + 18u) >> 5);  // cells j >= 14 use m / 2
  const double m = bits_to_double((ub & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  const double r = fma_(m, tab[2 * j], -1.0);
  const double de = (double)e;
  // -2 log1p(r) = r (-2 + r - 2/3 r^2 + 1/2 r^3 - 2/5 r^4 + 1/3 r^5 - 2/7 r^6 + 1/4 r^7 - 2/9 r^8)
#if defined(__CUDA_ARCH__)
  double p = __dadd_rn(__dmul_rn(r, -0x1.c71c7p-3), 0.25);  // (one immediate per instruction: DMUL + DADD instead of MOV, MOV, DFMA)
#else
  double p = r * -0x1.c71c7p-3 + 0.25;
#endif
  p = fma_(p, r, -0x1.24925p-2);
  p = fma_(p, r, mc.v[NC_L6]);
  p = fma_(p, r, mc.v[NC_L5]);
  p = fma_(p, r, 0.5);
  p = fma_(p, r, mc.v[NC_L3]);
  p = fma_(p, r, 1.0);
  p = fma_(p, r, -2.0);
  // -2 ln 2 = K_hi + K_lo, K_hi with 20 significant bits: de * K_hi is exact
  const double w = fma_(r, p, de * mc.v[NC_KLO]);
  return fma_(de, -0x1.62e43p+0, tab[2 * j + 1]) + w;
}

// (sin, cos)(2 pi u) for u = (k + 1/2) 2^-52, k = top 52 bits of b
MGMC_HD void sincos2pi_u52(uint64_t b, const NormalConsts &mc, const double *tab, double &sn, double &cs) {
  const uint32_t k = (uint32_t)(b >> 59);
  // 2 F + 1 for the 47 bits F below the cell index, minus 2^47: twice the offset from the cell centre, exact
  const double d = bits_to_double(0x4330000000000000ull | ((b >> 11) & 0x0000FFFFFFFFFFFFull) | 1ull) - 4644337115725824.0;
  const double y = d * mc.v[NC_ASCALE];
  const double z = y * y;
  double ps = fma_(z, 0x1.71de4p-19, mc.v[NC_S7]);
  ps = fma_(ps, z, mc.v[NC_S5]);
  ps = fma_(ps, z, mc.v[NC_S3]);
  const double sy = fma_(y * z, ps, y);  // sin y
  double pc = fma_(z, 0x1.a01ap-16, mc.v[NC_C6]);
  pc = fma_(pc, z, mc.v[NC_C4]);
  pc = fma_(pc, z, -0.5);
  const double cm = z * pc;  // cos y - 1
  const double S = tab[64 + 2 * k], C = tab[64 + 2 * k + 1];
  sn = fma_(C, sy, fma_(S, cm, S));
  cs = fma_(-S, sy, fma_(C, cm, C));
}

// Rounds: Philox4x32-R with R = 7, the smallest round count of the family that passes BigCrush (Salmon, Moraes, Dror,
// Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11, table 2: Philox4x32-7 is "Crush-resistant"; R = 10 is
// Random123's default for its safety margin).  The sweeps are instruction-issue bound and the generator is a third of
// their instructions: 7 rounds save 12 of the 149 instructions of a row iteration.  -DMGMC_PHILOX_ROUNDS=10 restores the
// default of Random123 / cuRAND (the oracle has the same switch and must be built with the same value).
#ifndef MGMC_PHILOX_ROUNDS
#define MGMC_PHILOX_ROUNDS 7
#endif
constexpr int kPhiloxRounds = MGMC_PHILOX_ROUNDS;

// Round keys of Philox4x32-R for a 64-bit seed: computed once on the host and passed by value in the
// kernel parameters, so that every round is 2 IMAD.WIDE + 2 LOP3 with the key read from the constant bank.
struct PhiloxKeys {
  uint32_t k[20];
};
inline PhiloxKeys philox_round_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {  // (room for 10 rounds; kPhiloxRounds of them are used)
    K.k[2 * r] = k0;
    K.k[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return K;
}

MGMC_HD void philox4x32(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, const PhiloxKeys &K) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < kPhiloxRounds; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
    c0 = n0;
    c1 = (uint32_t)p1;
    c2 = n2;
    c3 = (uint32_t)p0;
  }
}

// Box-Muller on two 64-bit words
MGMC_HD void box_muller(uint64_t a, uint64_t b, const NormalConsts &mc, const double *tab, double &z0, double &z1) {
  const double r = sqrt_(neg2log_u52(a, mc, tab));
  double s, c;
  sincos2pi_u52(b, mc, tab, s, c);
  z0 = r * c;
  z1 = r * s;
}

// two independent N(0,1) variates from one counter; tab = the 128-entry table above (shared, global or host memory)
MGMC_HD void normal_pair(const PhiloxKeys &K, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const NormalConsts &mc, const double *tab, double &z0,
                         double &z1) {
  philox4x32(c0, c1, c2, c3, K);
  box_muller((uint64_t)c0 | ((uint64_t)c1 << 32), (uint64_t)c2 | ((uint64_t)c3 << 32), mc, tab, z0, z1);
}

}  // namespace mgmc
