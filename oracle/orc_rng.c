/*
 * oracle/orc_rng.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the GSL generators behind NcmRNG
 * (numcosmo/ncm/core/ncm_rng.c:473,550 -> gsl_rng_mt19937 + gsl_rng_set;
 *  :696 gsl_rng_uniform, :713 gsl_rng_uniform_pos, :732 gsl_ran_flat,
 *  :752 gsl_ran_gaussian, :771 gsl_ran_ugaussian, :869 gsl_ran_beta,
 *  :907 gsl_ran_chisq).
 *
 * GSL (>= 2.8, meson.build:420) is NOT vendored under /root/reference and is
 * not installed in this image, so this file restates GSL's published
 * algorithms:  rng/mt.c (MT19937, Matsumoto & Nishimura 2002 initialisation),
 * randist/gauss.c (polar Box-Muller, second variate discarded),
 * randist/flat.c, randist/gamma.c (Marsaglia-Tsang), randist/gausszig.c
 * (Marsaglia-Tsang ziggurat, 128 levels, J. Voss), randist/chisq.c,
 * randist/beta.c.
 *
 * PARITY STATUS: MT19937 is pinned by the generator's published known-answer
 * (seed 5489 -> first output 3499211612, 10000th output 4123659995; checked in
 * tests/test_oracle_rng.py).  uniform / uniform_pos / flat / polar-gaussian are
 * simple closed forms over it.  The ziggurat tables (ytab/ktab/wtab) are
 * hard-coded 12-digit constants in GSL which cannot be reproduced from memory;
 * here they are REGENERATED from the ziggurat construction and rounded to 12
 * significant digits, so gsl_ran_gamma / gsl_ran_chisq / gsl_ran_beta streams
 * are "parity unpinned" (self-consistent between oracle and product, not
 * certified against a GSL build).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ncm_oracle.h"

#define MT_N 624
#define MT_M 397

static const unsigned long UPPER_MASK = 0x80000000UL;
static const unsigned long LOWER_MASK = 0x7fffffffUL;

/* rng/mt.c: mt_set() */
void
orc_rng_set (orc_rng *r, unsigned long s)
{
  int i;

  if (s == 0)
    s = 4357;

  r->mt[0] = s & 0xffffffffUL;

  for (i = 1; i < MT_N; i++)
  {
    r->mt[i]  = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long) i);
    r->mt[i] &= 0xffffffffUL;
  }

  r->mti = i;
}

/* rng/mt.c: mt_get() */
unsigned long
orc_rng_get (orc_rng *r)
{
  unsigned long k;
  unsigned long * const mt = r->mt;

#define MAGIC(y) (((y) & 0x1) ? 0x9908b0dfUL : 0)

  if (r->mti >= MT_N)
  {
    int kk;

    for (kk = 0; kk < MT_N - MT_M; kk++)
    {
      unsigned long y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);

      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ MAGIC (y);
    }

    for ( ; kk < MT_N - 1; kk++)
    {
      unsigned long y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);

      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ MAGIC (y);
    }

    {
      unsigned long y = (mt[MT_N - 1] & UPPER_MASK) | (mt[0] & LOWER_MASK);

      mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ MAGIC (y);
    }

    r->mti = 0;
  }

  k  = mt[r->mti];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);

  r->mti++;

  return k;
}

/* gsl_rng_uniform: mt_get_double = get / 4294967296.0  (ncm_rng.c:696) */
double
orc_rng_uniform (orc_rng *r)
{
  return orc_rng_get (r) / 4294967296.0;
}

/* gsl_rng_uniform_pos (ncm_rng.c:713) */
double
orc_rng_uniform_pos (orc_rng *r)
{
  double x;

  do {
    x = orc_rng_uniform (r);
  } while (x == 0);

  return x;
}

/* gsl_ran_flat (ncm_rng.c:732) */
double
orc_ran_flat (orc_rng *r, const double a, const double b)
{
  double u = orc_rng_uniform (r);

  return a * (1 - u) + b * u;
}

/* gsl_ran_gaussian: polar (Box-Muller) method (ncm_rng.c:752,771) */
double
orc_ran_gaussian (orc_rng *r, const double sigma)
{
  double x, y, r2;

  do {
    x = -1 + 2 * orc_rng_uniform_pos (r);
    y = -1 + 2 * orc_rng_uniform_pos (r);

    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);

  return sigma * y * sqrt (-2.0 * log (r2) / r2);
}

double
orc_ran_ugaussian (orc_rng *r)
{
  return orc_ran_gaussian (r, 1.0);
}

/* ---- ziggurat (randist/gausszig.c: Voss' variant, 128 levels).  GSL's tables rebuilt from their construction by
 * tools/gen_gausszig_tables.py (12-digit literals as in GSL's source; pinned on the entries of the GSL tables listed there) ---- */

#include "orc_gausszig_tables.h"
#define ZIG_R NCM_GAUSSZIG_PARAM_R
#define zig_ytab ncm_gausszig_ytab
#define zig_ktab ncm_gausszig_ktab
#define zig_wtab ncm_gausszig_wtab

double
orc_ran_gaussian_ziggurat (orc_rng *r, const double sigma)
{
  unsigned long i, j;
  int sign;
  double x, y;

  while (1)
  {
    /* mt19937: range = 0xffffffff, offset = 0 */
    unsigned long k = orc_rng_get (r);

    i = (k & 0xFF);
    j = (k >> 8) & 0xFFFFFF;

    sign = (i & 0x80) ? +1 : -1;
    i   &= 0x7f;

    x = j * zig_wtab[i];

    if (j < zig_ktab[i])
      break;

    if (i < 127)
    {
      double y0, y1, U1;

      y0 = zig_ytab[i];
      y1 = zig_ytab[i + 1];
      U1 = orc_rng_uniform (r);
      y  = y1 + (y0 - y1) * U1;
    }
    else
    {
      double U1, U2;

      U1 = 1.0 - orc_rng_uniform (r);
      U2 = orc_rng_uniform (r);
      x  = ZIG_R - log (U1) / ZIG_R;
      y  = exp (-ZIG_R * (x - 0.5 * ZIG_R)) * U2;
    }

    if (y < exp (-0.5 * x * x))
      break;
  }

  return sign * sigma * x;
}

/* randist/gamma.c: gsl_ran_gamma (Marsaglia & Tsang 2000) */
double
orc_ran_gamma (orc_rng *r, const double a, const double b)
{
  if (a < 1)
  {
    double u = orc_rng_uniform_pos (r);

    return orc_ran_gamma (r, 1.0 + a, b) * pow (u, 1.0 / a);
  }

  {
    double x, v, u;
    double d = a - 1.0 / 3.0;
    double c = (1.0 / 3.0) / sqrt (d);

    while (1)
    {
      do {
        x = orc_ran_gaussian_ziggurat (r, 1.0);
        v = 1.0 + c * x;
      } while (v <= 0);

      v = v * v * v;
      u = orc_rng_uniform_pos (r);

      if (u < 1 - 0.0331 * x * x * x * x)
        break;

      if (log (u) < 0.5 * x * x + d * (1 - v + log (v)))
        break;
    }

    return b * d * v;
  }
}

/* randist/chisq.c (ncm_rng.c:907) */
double
orc_ran_chisq (orc_rng *r, const double nu)
{
  return 2 * orc_ran_gamma (r, nu / 2, 1.0);
}

/* randist/beta.c (ncm_rng.c:869); Johnk's branch for a,b <= 1 */
double
orc_ran_beta (orc_rng *r, const double a, const double b)
{
  if ((a <= 1.0) && (b <= 1.0))
  {
    double U, V, X, Y;

    while (1)
    {
      U = orc_rng_uniform_pos (r);
      V = orc_rng_uniform_pos (r);
      X = pow (U, 1.0 / a);
      Y = pow (V, 1.0 / b);

      if ((X + Y) <= 1.0)
      {
        if (X + Y > 0)
        {
          return X / (X + Y);
        }
        else
        {
          double logX = log (U) / a;
          double logY = log (V) / b;
          double logM = logX > logY ? logX : logY;

          logX -= logM;
          logY -= logM;

          return exp (logX - log (exp (logX) + exp (logY)));
        }
      }
    }
  }
  else
  {
    double x1 = orc_ran_gamma (r, a, 1.0);
    double x2 = orc_ran_gamma (r, b, 1.0);

    return x1 / (x1 + x2);
  }
}
