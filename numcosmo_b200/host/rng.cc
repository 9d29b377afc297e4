// NcmRNG for the host mirror: GSL's default generator and the distributions the APES path draws from
// (numcosmo/ncm/core/ncm_rng.c:473,550,696,713,732,752,771,869,907).  GSL is not available in this image;
// the algorithms follow GSL's published sources (rng/mt.c, randist/{gauss,flat,gamma,gausszig,chisq,beta}.c).
// The proposal stream must be consumed in exactly the reference order (hard part a),
// which is why sampling stays on the host and only the affine map is offloaded.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "internal.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;

void mt_set(NcmRNG *r, unsigned long s) {
  if (s == 0) s = 4357;
  r->mt[0] = s & 0xffffffffUL;
  int i;
  for (i = 1; i < MT_N; i++) {
    r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long) i);
    r->mt[i] &= 0xffffffffUL;
  }
  r->mti = i;
}

inline unsigned long mt_get(NcmRNG *r) {
  unsigned long *const mt = r->mt;
  if (r->mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      const unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk]                = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
    }
    for (; kk < MT_N - 1; kk++) {
      const unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk]                = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
    }
    {
      const unsigned long y = (mt[MT_N - 1] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
      mt[MT_N - 1]          = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
    }
    r->mti = 0;
  }
  unsigned long k = mt[r->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  return k;
}

inline double uni(NcmRNG *r) { return mt_get(r) / 4294967296.0; }
inline double uni_pos(NcmRNG *r) {
  double x;
  do {
    x = uni(r);
  } while (x == 0);
  return x;
}

double gaussian_polar(NcmRNG *r, double sigma) {
  double x, y, r2;
  do {
    x  = -1 + 2 * uni_pos(r);
    y  = -1 + 2 * uni_pos(r);
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);
  return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

// gsl_ran_gaussian_ziggurat (randist/gausszig.c: Voss' variant of the Marsaglia-Tsang ziggurat, 128 levels).  GSL's three tables are
// rebuilt from their construction by tools/gen_gausszig_tables.py (the closure condition fixes R = 3.444286476761...; the entries of
// the GSL source it is pinned on are listed there) and included as the 12-digit literals GSL carries.
#include "gausszig_tables.h"
constexpr double ZIG_R = NCM_GAUSSZIG_PARAM_R;
const double *const zig_y = ncm_gausszig_ytab, *const zig_w = ncm_gausszig_wtab;
const unsigned long *const zig_k = ncm_gausszig_ktab;

double gaussian_zig(NcmRNG *r, double sigma) {
  unsigned long i, j;
  int sign;
  double x, y;
  while (true) {
    const unsigned long k = mt_get(r);
    i    = (k & 0xFF);
    j    = (k >> 8) & 0xFFFFFF;
    sign = (i & 0x80) ? +1 : -1;
    i &= 0x7f;
    x = j * zig_w[i];
    if (j < zig_k[i]) break;
    if (i < 127) {
      const double y0 = zig_y[i], y1 = zig_y[i + 1];
      const double U1 = uni(r);
      y               = y1 + (y0 - y1) * U1;
    } else {
      const double U1 = 1.0 - uni(r);
      const double U2 = uni(r);
      x               = ZIG_R - log(U1) / ZIG_R;
      y               = exp(-ZIG_R * (x - 0.5 * ZIG_R)) * U2;
    }
    if (y < exp(-0.5 * x * x)) break;
  }
  return sign * sigma * x;
}

double gamma_mt(NcmRNG *r, double a, double b) {
  if (a < 1) {
    const double u = uni_pos(r);
    return gamma_mt(r, 1.0 + a, b) * pow(u, 1.0 / a);
  }
  double x, v, u;
  const double d = a - 1.0 / 3.0;
  const double c = (1.0 / 3.0) / sqrt(d);
  while (true) {
    do {
      x = gaussian_zig(r, 1.0);
      v = 1.0 + c * x;
    } while (v <= 0);
    v = v * v * v;
    u = uni_pos(r);
    if (u < 1 - 0.0331 * x * x * x * x) break;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
  }
  return b * d * v;
}

}   // namespace

extern "C" {

NcmRNG *ncm_rng_new(const gchar *algo) {
  if (algo != nullptr && strcmp(algo, "mt19937") != 0) {
    ncm_b200_error("ncm_rng_new: only the GSL default generator `mt19937' is available, got `%s'.", algo);
    return nullptr;
  }
  NcmRNG *r = new NcmRNG;
  r->seed   = 0;
  mt_set(r, 0);
  return r;
}
NcmRNG *ncm_rng_seeded_new(const gchar *algo, gulong seed) {
  NcmRNG *r = ncm_rng_new(algo);
  if (r != nullptr) ncm_rng_set_seed(r, seed);
  return r;
}
void ncm_rng_free(NcmRNG *rng) { delete rng; }
void ncm_rng_clear(NcmRNG **rng) {
  if (rng != nullptr && *rng != nullptr) {
    delete *rng;
    *rng = nullptr;
  }
}
void ncm_rng_set_seed(NcmRNG *rng, gulong seed) {
  rng->seed = seed;
  mt_set(rng, seed);
}
gulong ncm_rng_get_seed(NcmRNG *rng) { return rng->seed; }
gulong ncm_rng_gen_ulong(NcmRNG *rng) { return mt_get(rng); }
gdouble ncm_rng_uniform01_gen(NcmRNG *rng) { return uni(rng); }
gdouble ncm_rng_uniform01_pos_gen(NcmRNG *rng) { return uni_pos(rng); }
gdouble ncm_rng_uniform_gen(NcmRNG *rng, const gdouble xl, const gdouble xu) {
  const double u = uni(rng);   // gsl_ran_flat
  return xl * (1 - u) + xu * u;
}
gdouble ncm_rng_gaussian_gen(NcmRNG *rng, const gdouble mu, const gdouble sigma) { return gaussian_polar(rng, sigma) + mu; }
gdouble ncm_rng_ugaussian_gen(NcmRNG *rng) { return gaussian_polar(rng, 1.0); }
gdouble ncm_rng_chisq_gen(NcmRNG *rng, const gdouble nu) { return 2 * gamma_mt(rng, nu / 2, 1.0); }
gdouble ncm_rng_beta_gen(NcmRNG *rng, const gdouble a, const gdouble b) {
  if ((a <= 1.0) && (b <= 1.0)) {
    while (true) {
      const double U = uni_pos(rng), V = uni_pos(rng);
      const double X = pow(U, 1.0 / a), Y = pow(V, 1.0 / b);
      if ((X + Y) <= 1.0) {
        if (X + Y > 0) return X / (X + Y);
        double logX = log(U) / a, logY = log(V) / b;
        const double logM = logX > logY ? logX : logY;
        logX -= logM;
        logY -= logM;
        return exp(logX - log(exp(logX) + exp(logY)));
      }
    }
  }
  const double x1 = gamma_mt(rng, a, 1.0);
  const double x2 = gamma_mt(rng, b, 1.0);
  return x1 / (x1 + x2);
}

}   // extern "C"
