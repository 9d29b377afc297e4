/*
 * oracle/orc_kernel.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of NcmStatsDistKernelGauss / NcmStatsDistKernelST:
 *   numcosmo/ncm/stats/ncm_stats_dist_kernel_gauss.c:193-355
 *   numcosmo/ncm/stats/ncm_stats_dist_kernel_st.c:225-414
 *   numcosmo/ncm/algebra/ncm_matrix.c:1157-1185 (cholesky_lndet)
 */
#include <math.h>
#include <stdlib.h>
#include "ncm_oracle.h"
#include "orc_blas.h"

#define ORC_LN2PI 1.8378770664093454835606594728112352797227949472755668 /* ncm_c_ln2pi () */
#define ORC_LNPI  1.1447298858494001741434273513530587116472948129153115 /* ncm_c_lnpi () */

/* _kernel_gauss.c:193-199 ; _kernel_st.c:225-238 */
double
orc_kernel_get_rot_bandwidth (const orc_kernel *k, const double n)
{
  const int d = k->d;

  if (k->kind == ORC_KERNEL_GAUSS)
  {
    return pow (4.0 / (n * (d + 2.0)), 1.0 / (d + 4.0));
  }
  else
  {
    const double nu = (k->nu >= 3.0) ? k->nu : 3.0;

    return pow (
      16.0 * ((nu - 2) * (nu - 2)) * (1.0 + d + nu) * (3.0 + d + nu) /
      ((2.0 + d) * (d + nu) * (2.0 + d + nu) * (d + 2.0 * nu) * (2.0 + d + 2.0 * nu) * n),
      1.0 / (d + 4.0));
  }
}

/* ncm_matrix.c:1157-1185 */
double
orc_cholesky_lndet (const double *U, int n, int ld)
{
  const double lb = 1.0e-200;
  const double ub = 1.0e+200;
  double detL     = 1.0;
  long exponent   = 0;
  int i;

  for (i = 0; i < n; i++)
  {
    const double Lii   = fabs (U[i * ld + i]);
    const double ndetL = detL * Lii;

    if ((ndetL < lb) || (ndetL > ub))
    {
      int exponent_i = 0;

      detL      = frexp (ndetL, &exponent_i);
      exponent += exponent_i;
    }
    else
    {
      detL = ndetL;
    }
  }

  return 2.0 * (log (detL) + exponent * M_LN2);
}

/* _kernel_gauss.c:201-207 ; _kernel_st.c:240-252 */
double
orc_kernel_get_lnnorm (const orc_kernel *k, const double *cov_decomp, int ld)
{
  const int d = k->d;

  if (k->kind == ORC_KERNEL_GAUSS)
  {
    return 0.5 * (d * ORC_LN2PI + orc_cholesky_lndet (cov_decomp, d, ld));
  }
  else
  {
    const double lg_lnnorm   = lgamma (k->nu / 2.0) - lgamma ((k->nu + d) / 2.0);
    const double chol_lnnorm = 0.5 * orc_cholesky_lndet (cov_decomp, d, ld);
    const double nc_lnnorm   = (d / 2.0) * (ORC_LNPI + log (k->nu));

    return lg_lnnorm + nc_lnnorm + chol_lnnorm;
  }
}

/* _kernel_gauss.c:209-213 ; _kernel_st.c:254-262 */
double
orc_kernel_eval_unnorm (const orc_kernel *k, const double chi2)
{
  if (k->kind == ORC_KERNEL_GAUSS)
    return exp (-0.5 * chi2);
  else
    return pow (1.0 + chi2 / k->nu, -0.5 * (k->nu + k->d));
}

/* _kernel_gauss.c:215-244 ; _kernel_st.c:264-293 (stride-aware) */
void
orc_kernel_eval_unnorm_vec (const orc_kernel *k, const double *chi2, int chi2_stride, double *Ku, int Ku_stride, int n)
{
  int i;

  for (i = 0; i < n; i++)
  {
    const double chi2_i = chi2[i * chi2_stride];

    Ku[i * Ku_stride] = orc_kernel_eval_unnorm (k, chi2_i);
  }
}

/* _kernel_gauss.c:246-289 ; _kernel_st.c:295-340 */
void
orc_kernel_eval_sum0_gamma_lambda (const orc_kernel *k, const double *chi2, const double *weights, const double *lnnorms, double *lnK, int n, double *gamma, double *lambda)
{
  const double kappa = -0.5 * (k->nu + k->d);
  double lnt_max     = -INFINITY;
  int i, i_max = 0;

  for (i = 0; i < n; i++)
  {
    const double chi2_i = chi2[i];
    const double w_i    = weights[i];
    const double lnu_i  = lnnorms[i];
    double lnt_i;

    if (k->kind == ORC_KERNEL_GAUSS)
      lnt_i = -0.5 * chi2_i - lnu_i + log (w_i);
    else
      lnt_i = kappa * log1p (chi2_i / k->nu) - lnu_i + log (w_i);

    if (lnt_i > lnt_max)
    {
      i_max   = i;
      lnt_max = lnt_i;
    }

    lnK[i] = lnt_i;
  }

  lambda[0] = 0.0;

  for (i = 0; i < i_max; i++)
    lambda[0] += exp (lnK[i] - lnt_max);

  for (i = i_max + 1; i < n; i++)
    lambda[0] += exp (lnK[i] - lnt_max);

  gamma[0] = lnt_max;
}

/* _kernel_gauss.c:291-333 ; _kernel_st.c:342-386 */
void
orc_kernel_eval_sum1_gamma_lambda (const orc_kernel *k, const double *chi2, const double *weights, double lnnorm, double *lnK, int n, double *gamma, double *lambda)
{
  const double kappa = -0.5 * (k->nu + k->d);
  double lnt_max     = -INFINITY;
  int i, i_max = 0;

  for (i = 0; i < n; i++)
  {
    const double chi2_i = chi2[i];
    const double w_i    = weights[i];
    double lnt_i;

    if (k->kind == ORC_KERNEL_GAUSS)
      lnt_i = -0.5 * chi2_i + log (w_i);
    else
      lnt_i = kappa * log1p (chi2_i / k->nu) + log (w_i);

    if (lnt_i > lnt_max)
    {
      i_max   = i;
      lnt_max = lnt_i;
    }

    lnK[i] = lnt_i;
  }

  lambda[0] = 0.0;

  for (i = 0; i < i_max; i++)
    lambda[0] += exp (lnK[i] - lnt_max);

  for (i = i_max + 1; i < n; i++)
    lambda[0] += exp (lnK[i] - lnt_max);

  gamma[0] = lnt_max - lnnorm;
}

/* _kernel_gauss.c:335-355 ; _kernel_st.c:388-414 */
void
orc_kernel_sample (const orc_kernel *k, const double *cov_decomp, int ld, const double href, const double *mu, double *x, orc_rng *rng)
{
  const int d = k->d;
  int i;

  for (i = 0; i < d; i++)
  {
    const double u_i = orc_ran_ugaussian (rng);

    x[i] = u_i * href;
  }

  /* gsl_blas_dtrmv (CblasUpper, CblasTrans, CblasNonUnit, cov_decomp, x) */
  scipy_cblas_dtrmv (OrcRowMajor, OrcUpper, OrcTrans, OrcNonUnit, d, cov_decomp, ld, x, 1);

  if (k->kind == ORC_KERNEL_ST)
  {
    const double chi_scale = sqrt (k->nu / orc_ran_chisq (rng, k->nu));

    for (i = 0; i < d; i++)
      x[i] *= chi_scale;
  }

  for (i = 0; i < d; i++)
    x[i] += mu[i];
}
