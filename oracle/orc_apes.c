/*
 * oracle/orc_apes.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the APES walker and of the ESMCMC accept loop around it:
 *   numcosmo/ncm/fit/ncm_fit_esmcmc_walker_apes.c:509-608 (set_sys: BOTH methods build
 *       NcmStatsDistVKDE objects), :646-691 (prepare_random_walk), :693-739 (sample),
 *       :741-817 (setup), :819-859 (transition_prob), :861-891 (step), :907-919 (prob_norm)
 *   numcosmo/ncm/fit/ncm_fit_esmcmc.c:2136-2148 (get_jumps), :2151-2232 (run_interval),
 *       :2235-2288 (run)
 *   numcosmo/ncm/core/ncm_util.c:645-716 (log_gaussian_integral)
 * and of the three synthetic targets (formulas only):
 *   numcosmo/ncm/data/ncm_data_rosenbrock.c:106-113, ncm_data_funnel.c:113-132,
 *   Gaussian with covariance from ncm_matrix.c:1687-1770 (fill_rand_cov).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "ncm_oracle.h"
#include "orc_blas.h"

#define ORC_LN2PI 1.8378770664093454835606594728112352797227949472755668
#define ERF_BOUND 1.0 /* ncm_util.c */

static double
now_s (void)
{
  struct timespec ts;

  clock_gettime (CLOCK_MONOTONIC, &ts);

  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ncm_util.c:645-693 */
static double
log_normal_gaussian_integral (const double xl, const double xu, double *sign)
{
  if (xl == xu)
    return -INFINITY;

  {
    const double sqrt_half = M_SQRT1_2;
    double ul, uu;

    if (xl < xu)
    {
      ul    = xl * sqrt_half;
      uu    = xu * sqrt_half;
      *sign = 1.0;
    }
    else
    {
      ul    = xu * sqrt_half;
      uu    = xl * sqrt_half;
      *sign = -1.0;
    }

    if (ul > ERF_BOUND)
    {
      const double val = 0.5 * (erfc (ul) - erfc (uu));

      return log (fabs (val));
    }
    else if (uu < -ERF_BOUND)
    {
      const double val = 0.5 * (erfc (-ul) - erfc (-uu));

      return log (fabs (val));
    }
    else if ((uu > ERF_BOUND) && (ul < -ERF_BOUND))
    {
      const double val = -0.5 * (erfc (uu) + erfc (-ul));

      return log1p (val);
    }
    else
    {
      const double val = 0.5 * (erf (uu) - erf (ul));

      return log (fabs (val));
    }
  }
}

/* ncm_util.c:709-716 */
double
orc_log_gaussian_integral (double xl, double xu, double mu, double sigma, double *sign)
{
  const double zl = (xl - mu) / sigma;
  const double zu = (xu - mu) / sigma;

  return log_normal_gaussian_integral (zl, zu, sign);
}

/* ncm_matrix.c:1687-1770 (fill_rand_cor + fill_rand_cov) */
void
orc_fill_rand_cov (double *cm, int n, double sigma_min, double sigma_max, double cor_level, orc_rng *rng)
{
  double *P = (double *) calloc ((size_t) n * n, sizeof (double));
  int k, i, l;

  for (i = 0; i < n * n; i++)
    cm[i] = 0.0;

  for (i = 0; i < n; i++)
    cm[i * n + i] = 1.0;

  for (k = 0; k < n - 1; k++)
  {
    for (i = k + 1; i < n; i++)
    {
      double p = (orc_ran_beta (rng, cor_level, cor_level) - 0.5) * 2.0;

      P[k * n + i] = p;

      for (l = k - 1; l >= 0; l--)
      {
        const double Pli = P[l * n + i];
        const double Plk = P[l * n + k];

        p = p * sqrt ((1.0 - Pli * Pli) * (1.0 - Plk * Plk)) + Pli * Plk;
      }

      cm[k * n + i] = p;
      cm[i * n + k] = p;
    }
  }

  for (k = 0; k < n; k++)
  {
    const double sigma_k = orc_ran_flat (rng, sigma_min, sigma_max);

    for (i = 0; i < n; i++)
      cm[i * n + k] *= sigma_k;

    for (i = 0; i < n; i++)
      cm[k * n + i] *= sigma_k;
  }

  free (P);
}

double
orc_target_m2lnL (const orc_target *t, const double *x)
{
  switch (t->kind)
  {
    case ORC_TARGET_MVND:
    {
      double v[64];
      double s = 0.0;
      int i, j;

      /* forward substitution with U^T (lower), chi2 = |U^-T (x - mu)|^2 */
      for (i = 0; i < t->d; i++)
      {
        double a = x[i] - t->mu[i];

        for (j = 0; j < i; j++)
          a -= t->cov_inv_U[j * t->d + i] * v[j];

        v[i] = a / t->cov_inv_U[i * t->d + i];
        s   += v[i] * v[i];
      }

      return s;
    }
    case ORC_TARGET_ROSENBROCK:
    {
      /* ncm_data_rosenbrock.c:106-113 */
      const double x1 = x[0], x2 = x[1];
      const double a  = x2 - x1 * x1, b = 1.0 - x1;

      return (100.0 * (a * a) + (b * b)) * 1.0e-1;
    }
    case ORC_TARGET_FUNNEL:
    {
      /* ncm_data_funnel.c:113-132: x[0] = nu, x[1..] = x_i */
      const double nu       = x[0];
      const double sigma_nu = exp (0.5 * nu);
      const int x_len       = t->d - 1;
      double m2lnL          = x_len * nu + (nu / 3.0) * (nu / 3.0);
      int i;

      for (i = 0; i < x_len; i++)
      {
        const double r = x[1 + i] / sigma_nu;

        m2lnL += r * r;
      }

      return m2lnL;
    }
    default:

      return NAN;
  }
}

typedef struct orc_apes_rw
{
  double *std, *lb, *ub;
} orc_apes_rw;

struct orc_apes
{
  int size, size_2, nparams;
  long n_lu_total, n_qr_total, n_nnls_total; /* instrumentation for the tests: NNLS solves of the run and how many left dposv */
  unsigned int exploration; /* walker_apes.c:142, 182: setup calls left in which the proposal density is left out of the acceptance */
  orc_sd *sd0, *sd1;
  double *thetastar, *m2lnp_star, *m2lnp_cur, *m2lnL_s0, *m2lnL_s1, *jumps;
  orc_apes_rw rw0, rw1;
  int use_interp, use_threads;
  double random_walk_prob;
  double timers[6];
};

orc_apes *
orc_apes_new (int nwalkers, int d, int sd_type, int kernel_kind, double nu, double over_smooth, int use_interp, double shrink, double random_walk_prob, double local_frac, int use_threads)
{
  orc_apes *a = (orc_apes *) calloc (1, sizeof (orc_apes));

  a->size    = nwalkers;
  a->size_2  = nwalkers / 2;
  a->nparams = d;
  a->sd0     = orc_sd_new (sd_type, kernel_kind, nu, d, ORC_CV_NONE);
  a->sd1     = orc_sd_new (sd_type, kernel_kind, nu, d, ORC_CV_NONE);

  orc_sd_set_over_smooth (a->sd0, over_smooth);
  orc_sd_set_over_smooth (a->sd1, over_smooth);
  orc_sd_set_shrink (a->sd0, shrink);
  orc_sd_set_shrink (a->sd1, shrink);
  orc_sd_set_use_threads (a->sd0, use_threads);
  orc_sd_set_use_threads (a->sd1, use_threads);
  orc_sd_set_local_frac (a->sd0, local_frac);
  orc_sd_set_local_frac (a->sd1, local_frac);

  a->use_interp       = use_interp;
  a->use_threads      = use_threads;
  a->random_walk_prob = random_walk_prob;

  a->thetastar  = (double *) calloc ((size_t) nwalkers * d, sizeof (double));
  a->m2lnp_star = (double *) calloc (nwalkers, sizeof (double));
  a->m2lnp_cur  = (double *) calloc (nwalkers, sizeof (double));
  a->m2lnL_s0   = (double *) calloc (nwalkers / 2, sizeof (double));
  a->m2lnL_s1   = (double *) calloc (nwalkers / 2, sizeof (double));
  a->jumps      = (double *) calloc (nwalkers, sizeof (double));
  a->rw0.std    = (double *) calloc (d, sizeof (double));
  a->rw0.lb     = (double *) calloc (d, sizeof (double));
  a->rw0.ub     = (double *) calloc (d, sizeof (double));
  a->rw1.std    = (double *) calloc (d, sizeof (double));
  a->rw1.lb     = (double *) calloc (d, sizeof (double));
  a->rw1.ub     = (double *) calloc (d, sizeof (double));

  return a;
}

/* ncm_fit_esmcmc_walker_apes.c:1442-1507: the covariance setters act on both halves; cov_fixed != NULL with ORC_COV_FIXED is
 * set_cov_fixed_from_mset's diag (scale^2) matrix */
/* ncm_fit_esmcmc_walker_apes.c:1523-1528 */
void
orc_apes_set_exploration (orc_apes *a, unsigned int exploration)
{
  a->exploration = exploration;
}

void
orc_apes_set_cov_type (orc_apes *a, int cov_type, const double *cov_fixed, int ld)
{
  orc_sd_set_cov_type (a->sd0, cov_type);
  orc_sd_set_cov_type (a->sd1, cov_type);

  if (cov_fixed != NULL)
  {
    orc_sd_set_cov_fixed (a->sd0, cov_fixed, ld);
    orc_sd_set_cov_fixed (a->sd1, cov_fixed, ld);
  }
}

void
orc_apes_free (orc_apes *a)
{
  orc_sd_free (a->sd0);
  orc_sd_free (a->sd1);
  free (a->thetastar);
  free (a->m2lnp_star);
  free (a->m2lnp_cur);
  free (a->m2lnL_s0);
  free (a->m2lnL_s1);
  free (a->jumps);
  free (a->rw0.std);
  free (a->rw0.lb);
  free (a->rw0.ub);
  free (a->rw1.std);
  free (a->rw1.lb);
  free (a->rw1.ub);
  free (a);
}

static int
valid_bounds (const orc_target *t, const double *x)
{
  int i;

  for (i = 0; i < t->d; i++)
  {
    if ((x[i] < t->lb[i]) || (x[i] > t->ub[i]))
      return 0;
  }

  return 1;
}

/* walker_apes.c:646-691 */
static void
prepare_random_walk (orc_apes *a, orc_sd *sd, orc_apes_rw *rw, const orc_target *t)
{
  if (a->random_walk_prob > 0.0)
  {
    const double *cov = orc_sd_peek_full_cov (sd);
    int i;

    for (i = 0; i < a->nparams; i++)
    {
      const double var = cov[i * a->nparams + i];

      rw->lb[i]  = t->lb[i];
      rw->ub[i]  = t->ub[i];
      rw->std[i] = sqrt (var) * 0.25;
    }
  }
}

/* walker_apes.c:693-714 */
static void
random_walk_sample (orc_apes *a, const orc_apes_rw *rw, const double *theta, double *thetastar, orc_rng *rng)
{
  int i;

  for (i = 0; i < a->nparams; i++)
  {
    const double lb      = rw->lb[i];
    const double ub      = rw->ub[i];
    const double std     = rw->std[i];
    const double theta_i = theta[i];
    double x;

    do {
      x = orc_ran_gaussian (rng, std) + theta_i; /* ncm_rng_gaussian_gen, ncm_rng.c:752 */
    } while ((x < lb) || (x > ub));

    thetastar[i] = x;
  }
}

/* walker_apes.c:716-739 */
static void
apes_sample (orc_apes *a, orc_sd *sd, const orc_target *t, const orc_apes_rw *rw, const double *theta, double *thetastar, orc_rng *rng)
{
  if (a->random_walk_prob == 0.0)
  {
    do {
      orc_sd_sample (sd, thetastar, rng);
    } while (!valid_bounds (t, thetastar));
  }
  else
  {
    do {
      if (orc_rng_uniform_pos (rng) < a->random_walk_prob)
        random_walk_sample (a, rw, theta, thetastar, rng);
      else
        orc_sd_sample (sd, thetastar, rng);
    } while (!valid_bounds (t, thetastar));
  }
}

/* walker_apes.c:819-859, with the density value supplied (m2lnp_sd = eval_m2lnp (sd, thetastar)) */
static double
transition_prob (orc_apes *a, const orc_apes_rw *rw, const double *theta, const double *thetastar, double m2lnp_sd)
{
  double m2lnp, sign;

  if (a->random_walk_prob > 0.0)
  {
    double m2lnp_rw = 0.0;
    int i;

    for (i = 0; i < a->nparams; i++)
    {
      const double lb          = rw->lb[i];
      const double ub          = rw->ub[i];
      const double std         = rw->std[i];
      const double theta_i     = theta[i];
      const double thetastar_i = thetastar[i];
      const double ln_norm     = 0.5 * ORC_LN2PI + log (std) + orc_log_gaussian_integral (lb, ub, theta_i, std, &sign);
      const double r           = (thetastar_i - theta_i) / std;

      m2lnp_rw += r * r + 2.0 * ln_norm;
    }

    m2lnp_rw += -2.0 * log (a->random_walk_prob);
    m2lnp_sd += -2.0 * log1p (-a->random_walk_prob);

    if (m2lnp_sd < m2lnp_rw)
      m2lnp = m2lnp_sd - 2.0 * log1p (exp (-0.5 * (m2lnp_rw - m2lnp_sd)));
    else
      m2lnp = m2lnp_rw - 2.0 * log1p (exp (-0.5 * (m2lnp_sd - m2lnp_rw)));
  }
  else
  {
    m2lnp = m2lnp_sd;
  }

  return m2lnp;
}

/* walker_apes.c:741-817 for one block: centres = the OTHER half */
static void
apes_setup_block (orc_apes *a, const orc_target *t, double *theta, double *m2lnL, int block, orc_rng *rng)
{
  const int d        = a->nparams;
  orc_sd *sd         = block == 0 ? a->sd0 : a->sd1;
  orc_apes_rw *rw    = block == 0 ? &a->rw0 : &a->rw1;
  double *m2lnL_s    = block == 0 ? a->m2lnL_s0 : a->m2lnL_s1;
  const int c0       = block == 0 ? a->size_2 : 0;           /* first centre */
  const int k0       = block == 0 ? 0 : a->size_2;           /* first walker updated */
  double tm[3], t0;
  int i;

  orc_sd_reset (sd);

  for (i = 0; i < a->size_2; i++)
  {
    m2lnL_s[i] = m2lnL[c0 + i];
    orc_sd_add_obs (sd, &theta[(size_t) (c0 + i) * d]);
  }

  if (a->use_interp)
  {
    orc_nnls_stats st;

    orc_sd_prepare_interp (sd, m2lnL_s, a->size_2);
    orc_sd_get_nnls_stats (sd, &st);
    a->n_nnls_total++;
    a->n_lu_total += st.n_lu;
    a->n_qr_total += st.n_qr;
  }
  else
    orc_sd_prepare (sd);

  orc_sd_get_timers (sd, tm);
  a->timers[0] += tm[0];

  if (a->use_interp)
  {
    a->timers[1] += tm[1];
    a->timers[2] += tm[2];
  }

  prepare_random_walk (a, sd, rw, t);

  t0 = now_s ();

  for (i = k0; i < k0 + a->size_2; i++)
    apes_sample (a, sd, t, rw, &theta[(size_t) i * d], &a->thetastar[(size_t) i * d], rng);

  a->timers[3] += now_s () - t0;
}

/* ncm_fit_esmcmc.c:2151-2232 + walker_apes.c:861-919 for walkers [k0, k0 + size_2) */
static void
apes_run_block (orc_apes *a, const orc_target *t, double *theta, double *m2lnL, int block, unsigned char *accepted, int nthreads)
{
  const int d     = a->nparams;
  orc_sd *sd      = block == 0 ? a->sd0 : a->sd1;
  orc_apes_rw *rw = block == 0 ? &a->rw0 : &a->rw1;
  const int k0    = block == 0 ? 0 : a->size_2;
  const int n     = a->size_2;
  double *q_star  = (double *) malloc (sizeof (double) * n);
  double *q_cur   = (double *) malloc (sizeof (double) * n);
  double t0       = now_s (), t1;
  int k;

  /* the two eval_m2lnp calls of _apes_step (walker_apes.c:872-873), one walker per thread */
  orc_sd_eval_m2lnp_batch (sd, &a->thetastar[(size_t) k0 * d], d, n, q_star, nthreads);
  orc_sd_eval_m2lnp_batch (sd, &theta[(size_t) k0 * d], d, n, q_cur, nthreads);

  t1            = now_s ();
  a->timers[4] += t1 - t0;

  for (k = k0; k < k0 + n; k++)
  {
    const double *theta_k   = &theta[(size_t) k * d];
    const double *thetastar = &a->thetastar[(size_t) k * d];
    const double jump       = a->jumps[k];
    double prob             = 0.0;

    a->m2lnp_star[k] = transition_prob (a, rw, theta_k, thetastar, q_star[k - k0]);
    a->m2lnp_cur[k]  = transition_prob (a, rw, thetastar, theta_k, q_cur[k - k0]);

    if (valid_bounds (t, thetastar))
    {
      const double m2lnL_star = orc_target_m2lnL (t, thetastar);

      if (isfinite (m2lnL_star))
      {
        /* _ncm_fit_esmcmc_walker_apes_prob_norm, walker_apes.c:908-919: 0 while exploring */
        const double lnq   = a->exploration ? 0.0 : -0.5 * (a->m2lnp_cur[k] - a->m2lnp_star[k]);
        const double m2lnq = -2.0 * lnq;
        const double m2lnp = m2lnL_star - m2lnL[k] + m2lnq;

        prob = exp (-0.5 * m2lnp);
        prob = (prob < 1.0) ? prob : 1.0;
      }

      if (jump < prob)
      {
        memcpy (&theta[(size_t) k * d], thetastar, sizeof (double) * d);
        m2lnL[k] = m2lnL_star;

        if (accepted != NULL)
          accepted[k] = 1;
      }
    }
  }

  a->timers[5] += now_s () - t1;

  free (q_star);
  free (q_cur);
}

/* ncm_fit_esmcmc.c:2235-2288 with ki = 0 (catalog pre-populated, numcosmo_py/sampling/apes.py:146-166) */
void
orc_apes_run (orc_apes *a, const orc_target *t, double *theta, double *m2lnL, int iters, orc_rng *rng, unsigned char *accepted, int nthreads)
{
  int it, k;

  for (it = 0; it < iters; it++)
  {
    unsigned char *acc = (accepted != NULL) ? &accepted[(size_t) it * a->size] : NULL;

    if (acc != NULL)
      memset (acc, 0, a->size);

    /* _ncm_fit_esmcmc_get_jumps (0, W) */
    for (k = 0; k < a->size; k++)
      a->jumps[k] = orc_rng_uniform (rng);

    /* every setup call counts the exploration phase down once it has drawn its proposals (walker_apes.c:815-816) */
    apes_setup_block (a, t, theta, m2lnL, 0, rng);
    if (a->exploration > 0)
      a->exploration--;
    apes_run_block (a, t, theta, m2lnL, 0, acc, nthreads);
    apes_setup_block (a, t, theta, m2lnL, 1, rng);
    if (a->exploration > 0)
      a->exploration--;
    apes_run_block (a, t, theta, m2lnL, 1, acc, nthreads);
  }
}

/* test instrumentation: {prepare_interp calls, dsysv fallbacks, dgels fallbacks} since the object was made */
void
orc_apes_get_fallback_counts (const orc_apes *a, long *n3)
{
  n3[0] = a->n_nnls_total;
  n3[1] = a->n_lu_total;
  n3[2] = a->n_qr_total;
}

void
orc_apes_get_timers (const orc_apes *a, double *t6)
{
  memcpy (t6, a->timers, sizeof (double) * 6);
}

const double *orc_apes_peek_thetastar (const orc_apes *a) { return a->thetastar; }
const double *orc_apes_peek_m2lnp_star (const orc_apes *a) { return a->m2lnp_star; }
const double *orc_apes_peek_m2lnp_cur (const orc_apes *a) { return a->m2lnp_cur; }
