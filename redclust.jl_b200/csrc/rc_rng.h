// rc_rng.h -- counter-based structured random stream (Philox-4x32-10) shared by the kernels and
// the CPU oracle.
//
// The reference draws from Julia's task-local Xoshiro stream in program order
// (/root/reference/src/mcmc.jl:109,130,154,379,404,469,524-525; src/utils.jl:4).  Raw-bit replay
// of that stream is impossible without Julia, so replay happens at the level of SEMANTIC draws:
// every draw site of the sampler has an address (chain, iteration, site, a, b) and the value at
// that address is a pure function of (seed, address).  The same addresses are used by the CUDA
// kernel and by the oracle, so both consume identical uniforms regardless of execution order --
// which also lets the kernel skip work the reference discards without
// desynchronising the stream.
#pragma once
#include "rc_math.h"

enum rc_site {
  RC_SITE_R_NORMAL = 1,   // a = rejection attempt          (mcmc.jl:109)
  RC_SITE_R_ACCEPT = 2,   //                                 (mcmc.jl:130)
  RC_SITE_P_GAMMA_A = 3,  // a = Marsaglia-Tsang attempt     (mcmc.jl:154, first Gamma)
  RC_SITE_P_GAMMA_B = 4,  // a = attempt                     (mcmc.jl:154, second Gamma)
  RC_SITE_P_BOOST = 5,    // a = 0/1 which gamma (shape < 1 boost)
  RC_SITE_SM_PAIR = 6,    // chaperones                      (mcmc.jl:379)
  RC_SITE_SM_LAUNCH = 7,  // a = position in S               (mcmc.jl:404)
  RC_SITE_SM_RGIBBS = 8,  // a = scan index, b = position in S (utils.jl:4 via mcmc.jl:337)
  RC_SITE_SM_ACCEPT = 9,  //                                 (mcmc.jl:469)
  RC_SITE_SCAN = 10,      // a = point i (0-based), b = candidate index / 2 (u0: even, u1: odd candidate) (utils.jl:4 via mcmc.jl:249)
  RC_SITE_INIT = 11       // a = 0: r ~ Gamma, 1: p ~ Beta   (mcmc.jl:524-525), iteration 0
};

struct rc_u4 { uint32_t x, y, z, w; };

RC_HD uint64_t rc_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

RC_HD rc_u4 rc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  rc_u4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3; return o;
}

// Per-chain key.
RC_HD uint64_t rc_chain_key(uint64_t seed, uint64_t chain) {
  return rc_splitmix64(seed ^ rc_splitmix64(chain + 0x632BE59BD9B4E019ULL));
}

struct rc_draw { double u0, u1; };   // two uniforms in [0,1), 53-bit (Julia rand() convention)

RC_HD rc_draw rc_draw2(uint64_t key, uint32_t iter, uint32_t site, uint32_t mh, uint32_t a, uint32_t b) {
  rc_u4 o = rc_philox4x32_10(b, a, (site & 0xffu) | (mh << 8), iter, (uint32_t)key, (uint32_t)(key >> 32));
  uint64_t w0 = ((uint64_t)o.y << 32) | o.x, w1 = ((uint64_t)o.w << 32) | o.z;
  rc_draw d;
  d.u0 = (double)(w0 >> 11) * 0x1p-53;
  d.u1 = (double)(w1 >> 11) * 0x1p-53;
  return d;
}
RC_HD double rc_draw1(uint64_t key, uint32_t iter, uint32_t site, uint32_t mh, uint32_t a, uint32_t b) {
  return rc_draw2(key, iter, site, mh, a, b).u0;
}

// uniform on {1..n}  (semantic equivalent of Julia rand(1:n))
RC_HD int64_t rc_randint(double u, int64_t n) {
  int64_t k = (int64_t)(u * (double)n);
  if (k >= n) k = n - 1;
  return k + 1;
}
// strictly inside (0,1) for quantile transforms
RC_HD double rc_open01(double u) { return u + 0x1p-54; }

// Gamma(shape, 1) by Marsaglia-Tsang; attempts are addressed by `a` so the stream is structured.
RC_HD double rc_gamma_mt(double shape, uint64_t key, uint32_t iter, uint32_t site, uint32_t boost_a) {
  double k = shape;
  double boost = 1.0;
  if (k < 1.0) {
    double ub = rc_draw1(key, iter, RC_SITE_P_BOOST, 0, boost_a, site);
    boost = rc_exp(rc_log(ub) / k);
    k += 1.0;
  }
  double d = k - 1.0 / 3.0;
  double c = 1.0 / sqrt(9.0 * d);
  for (uint32_t att = 0; att < 100000u; ++att) {
    rc_draw dr = rc_draw2(key, iter, site, 0, att, 0);
    double z = rc_norminv(rc_open01(dr.u0));
    double t = 1.0 + c * z;
    if (t <= 0.0) continue;
    double v = t * t * t;
    if (rc_log(dr.u1) < 0.5 * z * z + d - d * v + d * rc_log(v)) return d * v * boost;
  }
  return d * boost;
}
RC_HD double rc_beta(double a, double b, uint64_t key, uint32_t iter) {
  double x = rc_gamma_mt(a, key, iter, RC_SITE_P_GAMMA_A, 0);
  double y = rc_gamma_mt(b, key, iter, RC_SITE_P_GAMMA_B, 1);
  return x / (x + y);
}
// r ~ Gamma(eta, 1/sigma), p ~ Beta(u, v): the default initial state of runsampler
// (/root/reference/src/mcmc.jl:524-525), drawn at iteration 0 of the chain's stream.
RC_HD void rc_init_rp_draw(double eta, double sigma, double u, double v, uint64_t seed, uint64_t chain,
                           double* r, double* p) {
  const uint64_t key = rc_chain_key(seed, chain);
  *r = rc_gamma_mt(eta, key, 0, RC_SITE_INIT, 2) * (1 / sigma);
  *p = rc_beta(u, v, key, 0);
}
