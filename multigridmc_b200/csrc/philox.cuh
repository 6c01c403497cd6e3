// Counter-based noise for the Gibbs sweeps: Philox4x32-10 (Salmon et al., SC'11) generated in
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
// The transcendental part is hand-written (fdlibm-style kernels, ~1 ulp) because the sweeps are
// otherwise instruction-issue bound: CUDA's log / sincospi / sqrt carry special-case paths that a
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

// Polynomial coefficients of the log / sin / cos kernels.  On the device they live in constant memory
// so that every DFMA reads its coefficient straight from the constant bank (no UMOV pairs per use).
#define MGMC_MATH_TABLE                                                                                                          \
  {6.93147180369123816490e-01, 1.90821492927058770002e-10, /* ln2_hi, ln2_lo */                                                 \
   6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01, /* Lg1..Lg4 */      \
   1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,                           /* Lg5..Lg7 */      \
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, /* S1..S3 */                           \
   2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,  /* S4..S6 */                           \
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,  /* C1..C3 */                           \
   -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11, /* C4..C6 */                           \
   3.14159265358979311600e+00, 1.22464679914735317723e-16, 6755399441055744.0}           /* pi_hi, pi_lo, 1.5 * 2^52 */
#if defined(__CUDACC__)
static __constant__ double kMathDev[24] = MGMC_MATH_TABLE;
#endif
static const double kMathHost[24] = MGMC_MATH_TABLE;
#if defined(__CUDA_ARCH__)
#define MC(i) kMathDev[i]
#else
#define MC(i) kMathHost[i]
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
// 1/d for d in [1.4, 3.5]: float seed + 2 Newton steps (full double precision)
MGMC_HD double rcp_(double d) {
#if defined(__CUDA_ARCH__)
  float rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)d));  // MUFU.RCP, 1 ulp: 2 Newton steps reach fp64
  double r = (double)rf;
#else
  double r = (double)(1.0f / (float)d);
#endif
  r = fma_(r, fma_(-d, r, 1.0), r);
  r = fma_(r, fma_(-d, r, 1.0), r);
  return r;
}
// sqrt(t) for t in (0, 80]: float rsqrt seed + 2 Newton steps on 1/sqrt + 1 correction on sqrt
MGMC_HD double sqrt_(double t) {
#if defined(__CUDA_ARCH__)
  float yf;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)t));  // MUFU.RSQ
  double y = (double)yf;
#else
  double y = (double)(1.0f / std::sqrt((float)t));
#endif
  const double h = 0.5 * t;
  y = y * fma_(-h * y, y, 1.5);
  y = y * fma_(-h * y, y, 1.5);
  double r = t * y;
  r = fma_(0.5 * y, fma_(-r, r, t), r);
  return r;
}

// -2 ln(u) for a double u in (0, 1)  (fdlibm __ieee754_log kernel, error < 1 ulp)
MGMC_HD double neg2log_(double u) {
  const double ln2_hi = MC(0), ln2_lo = MC(1);
  const double Lg1 = MC(2), Lg2 = MC(3), Lg3 = MC(4), Lg4 = MC(5), Lg5 = MC(6), Lg6 = MC(7), Lg7 = MC(8);
  uint64_t b = double_to_bits(u);
  int hx = (int)(b >> 32);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int i = (hx + 0x95f64) & 0x100000;  // 1 if mantissa > sqrt(2)
  b = (b & 0x000fffffffffffffull) | ((uint64_t)(0x3ff00000 ^ i) << 32);
  k += (i >> 20);
  const double f = bits_to_double(b) - 1.0;
  const double s = f * rcp_(2.0 + f);
  const double dk = (double)k;
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * fma_(w, fma_(w, Lg6, Lg4), Lg2);
  const double t2 = z * fma_(w, fma_(w, fma_(w, Lg7, Lg5), Lg3), Lg1);
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  const double lg = fma_(dk, ln2_hi, -((hfsq - fma_(s, hfsq + R, dk * ln2_lo)) - f));
  return -2.0 * lg;
}

// (sin, cos)(pi x) for x in (0, 2): octant reduction + fdlibm __kernel_sin / __kernel_cos
MGMC_HD void sincospi_(double x, double &sn, double &cs) {
  const double S1 = MC(9), S2 = MC(10), S3 = MC(11), S4 = MC(12), S5 = MC(13), S6 = MC(14);
  const double C1 = MC(15), C2 = MC(16), C3 = MC(17), C4 = MC(18), C5 = MC(19), C6 = MC(20);
  // n = nearest integer to 2x (0..4), r = x - n/2 in [-1/4, 1/4] exactly
  const double two52 = MC(23);  // 1.5 * 2^52: adding it rounds to nearest integer
  const double tn = fma_(2.0, x, two52);
  const int n = (int)(uint32_t)double_to_bits(tn);
  const double r = fma_(-0.5, tn - two52, x);
  const double y = fma_(r, MC(21), r * MC(22));
  const double z = y * y;
  const double ps = fma_(z, fma_(z, fma_(z, fma_(z, fma_(z, S6, S5), S4), S3), S2), S1);
  const double s0 = fma_(y * z, ps, y);
  const double pc = fma_(z, fma_(z, fma_(z, fma_(z, fma_(z, C6, C5), C4), C3), C2), C1);
  const double c0 = fma_(z * z, pc, fma_(-0.5, z, 1.0));
  const double a = (n & 1) ? c0 : s0;  // |sin|-branch
  const double b = (n & 1) ? s0 : c0;  // |cos|-branch
  sn = (n & 2) ? -a : a;
  cs = ((n + 1) & 2) ? -b : b;
}

// Round keys of Philox4x32-10 for a 64-bit seed: computed once on the host and passed by value in the
// kernel parameters, so that every round is 2 IMAD.WIDE + 2 LOP3 with the key read from the constant bank.
struct PhiloxKeys {
  uint32_t k[20];
};
inline PhiloxKeys philox_round_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    K.k[2 * r] = k0;
    K.k[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return K;
}

MGMC_HD void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, const PhiloxKeys &K) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
    c0 = n0;
    c1 = (uint32_t)p1;
    c2 = n2;
    c3 = (uint32_t)p0;
  }
}

// uniform in (0,1) with 52 random bits: (k + 1/2) 2^-52, k = top 52 bits
MGMC_HD double uniform52(uint64_t a) { return bits_to_double(0x3FF0000000000000ull | (a >> 12)) - (1.0 - 1.1102230246251565e-16); }

// Box-Muller on two 64-bit words
MGMC_HD void box_muller(uint64_t a, uint64_t b, double &z0, double &z1) {
  const double u1 = uniform52(a), u2 = uniform52(b);
  const double r = sqrt_(neg2log_(u1));
  double s, c;
  sincospi_(2.0 * u2, s, c);
  z0 = r * c;
  z1 = r * s;
}

// two independent N(0,1) variates from one counter
MGMC_HD void normal_pair(const PhiloxKeys &K, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double &z0, double &z1) {
  philox4x32_10(c0, c1, c2, c3, K);
  box_muller((uint64_t)c0 | ((uint64_t)c1 << 32), (uint64_t)c2 | ((uint64_t)c3 << 32), z0, z1);
}

}  // namespace mgmc
