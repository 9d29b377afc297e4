/*
 * oracle/orc_optim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The two derivative-free optimisers the cross-validation modes of NcmStatsDist drive
 * (numcosmo/ncm/stats/ncm_stats_dist.c:660-701 and :1018-1072):
 *
 *  - orc_nmsimplex2_*: GSL's gsl_multimin_fminimizer_nmsimplex2 (GSL >= 2.8, multimin/simplex2.c; GSL is a
 *    third-party dependency that is NOT under /root/reference).  Restated from the
 *    published algorithm: Nelder-Mead with the O(N) centre / size updates, coefficients -1 (reflection),
 *    -2 (expansion), 0.5 (contraction), shrink about the best corner; size = rms distance of the corners
 *    from the centre.  PARITY UNPINNED against a GSL build (none in this image).
 *
 *  - orc_lm_dif: levmar's dlevmar_dif (numcosmo/external/levmar/lm_core.c:436-851, forward-difference
 *    Jacobian misc_core.c:137-172, blocked squared norm misc_core.c:722-808, linear solve Axb_core.c:1141-1283)
 *    restated.  This one IS pinned: oracle/Makefile compiles the reference's own levmar sources into
 *    oracle/_ref/liblevmar_ref.so and tests/test_oracle_cv.py compares the two iterate by iterate.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include "ncm_oracle.h"

/* ------------------------------------------------------------------------------------------------ */
/* nmsimplex2                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

struct orc_nmsimplex2
{
  int n;          /* dimension; P = n + 1 corners */
  double *x1;     /* P x n corner points */
  double *y1;     /* P function values */
  double *ws1, *ws2;
  double *center, *delta, *xmc;
  double S2;
  unsigned long count;
  /* the fminimizer shell */
  double *x;
  double fval, size;
  orc_fmin_fn f;
  void *params;
};

static double
vnrm2 (const double *v, int n)
{
  /* BLAS dnrm2 semantics (scaled); for n = 1 this is |v| exactly */
  double scale = 0.0, ssq = 1.0;
  int i;

  if (n == 1)
    return fabs (v[0]);

  for (i = 0; i < n; i++)
  {
    if (v[i] != 0.0)
    {
      const double a = fabs (v[i]);

      if (scale < a)
      {
        ssq   = 1.0 + ssq * (scale / a) * (scale / a);
        scale = a;
      }
      else
      {
        ssq += (a / scale) * (a / scale);
      }
    }
  }

  return scale * sqrt (ssq);
}

static void
nm_compute_center (orc_nmsimplex2 *s, double *center)
{
  const int n = s->n, P = n + 1;
  int i, j;

  for (j = 0; j < n; j++)
    center[j] = 0.0;

  for (i = 0; i < P; i++)
    for (j = 0; j < n; j++)
      center[j] += 1.0 * s->x1[i * n + j];

  {
    const double alpha = 1.0 / P;

    for (j = 0; j < n; j++)
      center[j] *= alpha;
  }
}

static double
nm_compute_size (orc_nmsimplex2 *s, const double *center)
{
  const int n = s->n, P = n + 1;
  double ss   = 0.0;
  int i, j;

  for (i = 0; i < P; i++)
  {
    double t;

    for (j = 0; j < n; j++)
      s->ws1[j] = s->x1[i * n + j] + (-1.0) * center[j];

    t   = vnrm2 (s->ws1, n);
    ss += t * t;
  }

  s->S2 = ss / P;

  return sqrt (ss / P);
}

static double
nm_try_corner_move (const double coeff, orc_nmsimplex2 *s, int corner, double *xc)
{
  const int n = s->n, P = n + 1;
  const double alpha = (1 - coeff) * P / (P - 1.0);
  const double beta  = (P * coeff - 1.0) / (P - 1.0);
  int j;

  for (j = 0; j < n; j++)
    xc[j] = s->center[j];

  for (j = 0; j < n; j++)
    xc[j] *= alpha;

  for (j = 0; j < n; j++)
    xc[j] += beta * s->x1[corner * n + j];

  return s->f (xc, n, s->params);
}

static void
nm_update_point (orc_nmsimplex2 *s, int i, const double *x, double val)
{
  const int n = s->n, P = n + 1;
  double *x_orig = &s->x1[i * n];
  int j;

  for (j = 0; j < n; j++)
    s->delta[j] = x[j] + (-1.0) * x_orig[j];

  for (j = 0; j < n; j++)
    s->xmc[j] = x_orig[j] + (-1.0) * s->center[j];

  {
    const double d = vnrm2 (s->delta, n);
    double xmcd    = 0.0;

    for (j = 0; j < n; j++)
      xmcd += s->xmc[j] * s->delta[j];

    s->S2 += (2.0 / P) * xmcd + ((P - 1.0) / P) * (d * d / P);
  }

  {
    const double alpha = 1.0 / P;

    for (j = 0; j < n; j++)
      s->center[j] += (-alpha) * x_orig[j];

    for (j = 0; j < n; j++)
      s->center[j] += alpha * x[j];
  }

  for (j = 0; j < n; j++)
    x_orig[j] = x[j];

  s->y1[i] = val;
}

static int
nm_contract_by_best (orc_nmsimplex2 *s, int best, double *xc)
{
  const int n = s->n, P = n + 1;
  int status  = 0;
  int i, j;

  for (i = 0; i < P; i++)
  {
    if (i != best)
    {
      double newval;

      for (j = 0; j < n; j++)
        s->x1[i * n + j] = 0.5 * (s->x1[i * n + j] + s->x1[best * n + j]);

      for (j = 0; j < n; j++)
        xc[j] = s->x1[i * n + j];

      newval   = s->f (xc, n, s->params);
      s->y1[i] = newval;

      if (!isfinite (newval))
        status = 1; /* GSL_EBADFUNC */
    }
  }

  nm_compute_center (s, s->center);
  nm_compute_size (s, s->center);

  return status;
}

orc_nmsimplex2 *
orc_nmsimplex2_new (int n)
{
  orc_nmsimplex2 *s = (orc_nmsimplex2 *) calloc (1, sizeof (orc_nmsimplex2));
  const int P       = n + 1;

  s->n      = n;
  s->x1     = (double *) calloc ((size_t) P * n, sizeof (double));
  s->y1     = (double *) calloc (P, sizeof (double));
  s->ws1    = (double *) calloc (n, sizeof (double));
  s->ws2    = (double *) calloc (n, sizeof (double));
  s->center = (double *) calloc (n, sizeof (double));
  s->delta  = (double *) calloc (n, sizeof (double));
  s->xmc    = (double *) calloc (n, sizeof (double));
  s->x      = (double *) calloc (n, sizeof (double));
  s->count  = 0;

  return s;
}

void
orc_nmsimplex2_free (orc_nmsimplex2 *s)
{
  if (s == NULL)
    return;

  free (s->x1);
  free (s->y1);
  free (s->ws1);
  free (s->ws2);
  free (s->center);
  free (s->delta);
  free (s->xmc);
  free (s->x);
  free (s);
}

/* gsl_multimin_fminimizer_set + nmsimplex_set: corner 0 = x, corner i + 1 = x + step_i e_i */
int
orc_nmsimplex2_set (orc_nmsimplex2 *s, orc_fmin_fn f, void *params, const double *x, const double *step_size)
{
  const int n = s->n;
  double *xtemp = s->ws1;
  double val;
  int i, j;

  s->f      = f;
  s->params = params;

  for (j = 0; j < n; j++)
    s->x[j] = x[j];

  val = f (s->x, n, params);

  if (!isfinite (val))
    return 1;

  for (j = 0; j < n; j++)
    s->x1[j] = s->x[j];

  s->y1[0] = val;

  for (i = 0; i < n; i++)
  {
    for (j = 0; j < n; j++)
      xtemp[j] = s->x[j];

    xtemp[i] = s->x[i] + step_size[i];
    val      = f (xtemp, n, params);

    if (!isfinite (val))
      return 1;

    for (j = 0; j < n; j++)
      s->x1[(i + 1) * n + j] = xtemp[j];

    s->y1[i + 1] = val;
  }

  nm_compute_center (s, s->center);
  s->size = nm_compute_size (s, s->center);
  s->count++;

  return 0;
}

int
orc_nmsimplex2_iterate (orc_nmsimplex2 *s)
{
  const int n = s->n, P = n + 1;
  double *xc = s->ws1, *xc2 = s->ws2;
  double *y1 = s->y1;
  int hi, s_hi, lo, i, j;
  double dhi, ds_hi, dlo, val, val2;

  /* highest, second highest and lowest corner (with GSL's initialisation: second highest starts at corner 1) */
  dhi = dlo = y1[0];
  hi  = 0;
  lo  = 0;

  ds_hi = y1[1];
  s_hi  = 1;

  for (i = 1; i < P; i++)
  {
    val = y1[i];

    if (val < dlo)
    {
      dlo = val;
      lo  = i;
    }
    else if (val > dhi)
    {
      ds_hi = dhi;
      s_hi  = hi;
      dhi   = val;
      hi    = i;
    }
    else if (val > ds_hi)
    {
      ds_hi = val;
      s_hi  = i;
    }
  }

  val = nm_try_corner_move (-1.0, s, hi, xc);

  if (isfinite (val) && (val < y1[lo]))
  {
    val2 = nm_try_corner_move (-2.0, s, hi, xc2);

    if (isfinite (val2) && (val2 < y1[lo]))
      nm_update_point (s, hi, xc2, val2);
    else
      nm_update_point (s, hi, xc, val);
  }
  else if (!isfinite (val) || (val > y1[s_hi]))
  {
    if (isfinite (val) && (val <= y1[hi]))
      nm_update_point (s, hi, xc, val);

    val2 = nm_try_corner_move (0.5, s, hi, xc2);

    if (isfinite (val2) && (val2 <= y1[hi]))
    {
      nm_update_point (s, hi, xc2, val2);
    }
    else
    {
      if (nm_contract_by_best (s, lo, xc) != 0)
        return 1; /* GSL_EFAILED "contraction failed" */
    }
  }
  else
  {
    nm_update_point (s, hi, xc, val);
  }

  /* lowest corner becomes x (gsl_vector_min_index: first minimum) */
  lo = 0;

  for (i = 1; i < P; i++)
    if (y1[i] < y1[lo])
      lo = i;

  for (j = 0; j < n; j++)
    s->x[j] = s->x1[lo * n + j];

  s->fval = y1[lo];

  if (s->S2 > 0)
    s->size = sqrt (s->S2);
  else
    s->size = nm_compute_size (s, s->center);

  return 0;
}

const double *orc_nmsimplex2_x (const orc_nmsimplex2 *s) { return s->x; }
double orc_nmsimplex2_fval (const orc_nmsimplex2 *s) { return s->fval; }
double orc_nmsimplex2_size (const orc_nmsimplex2 *s) { return s->size; }

/* the loop of _ncm_stats_dist_minimize_obj (ncm_stats_dist.c:681-692): returns the iteration count */
int
orc_nmsimplex2_minimize (orc_nmsimplex2 *s, orc_fmin_fn f, void *params, const double *x0, const double *step, double size_tol, int max_iter)
{
  int iter = 0;

  if (orc_nmsimplex2_set (s, f, params, x0, step) != 0)
    return -1;

  for (;;)
  {
    iter++;

    if (orc_nmsimplex2_iterate (s) != 0)
      break;

    /* gsl_multimin_test_size */
    if (s->size < size_tol)
      break;

    if (!(iter < max_iter))
      break;
  }

  return iter;
}

/* ------------------------------------------------------------------------------------------------ */
/* Levenberg-Marquardt with a finite-difference (secant-updated) Jacobian: dlevmar_dif               */
/* ------------------------------------------------------------------------------------------------ */

/* e = x - y (or -y when x == NULL), returns ||e||^2 accumulated in the four interleaved partial sums of
 * misc_core.c:722-808 (blocks of 8 walked downwards, then the tail) */
static double
lm_l2nrmxmy (double *e, const double *x, const double *y, int n)
{
  const int blockn = (n >> 3) << 3;
  double sum[4]    = {0.0, 0.0, 0.0, 0.0};
  int i, k;

  for (i = blockn - 1; i > 0; i -= 8)
  {
    for (k = 0; k < 8; k++)
    {
      const int j = i - k;

      e[j]        = (x != NULL) ? x[j] - y[j] : -y[j];
      sum[k & 3] += e[j] * e[j];
    }
  }

  for (i = blockn; i < n; i++)
  {
    const int c = n - i; /* the switch label this element is handled under: 7 .. 1 */

    e[i]              = (x != NULL) ? x[i] - y[i] : -y[i];
    sum[(7 - c) & 3] += e[i] * e[i];
  }

  return sum[0] + sum[1] + sum[2] + sum[3];
}

/* m x m solve: row-scaled partial-pivoting LU (Axb_core.c:1141-1283); A, B untouched; returns 0 when singular */
static int
lm_solve (const double *A, const double *B, double *x, int m)
{
  double *a    = (double *) malloc (sizeof (double) * ((size_t) m * m + m));
  double *work = a + (size_t) m * m;
  int *idx     = (int *) malloc (sizeof (int) * m);
  int i, j, k, maxi = -1;
  double max, sum, tmp;

  memcpy (a, A, sizeof (double) * (size_t) m * m);
  memcpy (x, B, sizeof (double) * m);

  for (i = 0; i < m; i++)
  {
    max = 0.0;

    for (j = 0; j < m; j++)
      if ((tmp = fabs (a[i * m + j])) > max)
        max = tmp;

    if (max == 0.0)
    {
      free (a);
      free (idx);

      return 0;
    }

    work[i] = 1.0 / max;
  }

  for (j = 0; j < m; j++)
  {
    for (i = 0; i < j; i++)
    {
      sum = a[i * m + j];

      for (k = 0; k < i; k++)
        sum -= a[i * m + k] * a[k * m + j];

      a[i * m + j] = sum;
    }

    max = 0.0;

    for (i = j; i < m; i++)
    {
      sum = a[i * m + j];

      for (k = 0; k < j; k++)
        sum -= a[i * m + k] * a[k * m + j];

      a[i * m + j] = sum;

      if ((tmp = work[i] * fabs (sum)) >= max)
      {
        max  = tmp;
        maxi = i;
      }
    }

    if (j != maxi)
    {
      for (k = 0; k < m; k++)
      {
        tmp             = a[maxi * m + k];
        a[maxi * m + k] = a[j * m + k];
        a[j * m + k]    = tmp;
      }

      work[maxi] = work[j];
    }

    idx[j] = maxi;

    if (a[j * m + j] == 0.0)
      a[j * m + j] = DBL_EPSILON;

    if (j != m - 1)
    {
      tmp = 1.0 / a[j * m + j];

      for (i = j + 1; i < m; i++)
        a[i * m + j] *= tmp;
    }
  }

  for (i = k = 0; i < m; i++)
  {
    j    = idx[i];
    sum  = x[j];
    x[j] = x[i];

    if (k != 0)
    {
      for (j = k - 1; j < i; j++)
        sum -= a[i * m + j] * x[j];
    }
    else if (sum != 0.0)
    {
      k = i + 1;
    }

    x[i] = sum;
  }

  for (i = m - 1; i >= 0; i--)
  {
    sum = x[i];

    for (j = i + 1; j < m; j++)
      sum -= a[i * m + j] * x[j];

    x[i] = sum / a[i * m + i];
  }

  free (a);
  free (idx);

  return 1;
}

#define ORC_LM_BLOCKSZ_SQ (32 * 32) /* misc.h __BLOCKSZ__SQ: below it J^T J is accumulated row by row, downwards */

int
orc_lm_dif (orc_lm_fn func, double *p, const double *x, int m, int n, int itmax, const double *opts, double *info, void *adata)
{
  const int nm = n * m;
  const int K  = (m >= 10) ? m : 10;
  double *work, *e, *hx, *jacTe, *jac, *jacTjac, *Dp, *diag, *pDp, *wrk, *wrk2;
  double mu = 0.0, tmp, p_eL2, jacTe_inf = 0.0, pDp_eL2, p_L2 = 0.0, Dp_L2 = DBL_MAX, dF, dL, init_p_eL2;
  double tau, eps1, eps2, eps2_sq, eps3, delta;
  int nu, nu2, stop = 0, nfev, njap = 0, nlss = 0, updjac = 0, updp = 1, newjac = 0, using_ffdif = 1;
  int i, j, k, l;

  if (n < m)
    return -1;

  if (opts != NULL)
  {
    tau     = opts[0];
    eps1    = opts[1];
    eps2    = opts[2];
    eps2_sq = opts[2] * opts[2];
    eps3    = opts[3];
    delta   = opts[4];

    if (delta < 0.0)
    {
      delta       = -delta;
      using_ffdif = 0;
    }
  }
  else
  {
    tau     = 1e-3;  /* LM_INIT_MU */
    eps1    = 1e-17; /* LM_STOP_THRESH */
    eps2    = 1e-17;
    eps2_sq = 1e-17 * 1e-17;
    eps3    = 1e-17;
    delta   = 1e-6; /* LM_DIFF_DELTA */
  }

  work    = (double *) malloc (sizeof (double) * ((size_t) 4 * n + 4 * m + (size_t) n * m + (size_t) m * m));
  e       = work;
  hx      = e + n;
  jacTe   = hx + n;
  jac     = jacTe + m;
  jacTjac = jac + nm;
  Dp      = jacTjac + m * m;
  diag    = Dp + m;
  pDp     = diag + m;
  wrk     = pDp + m;
  wrk2    = wrk + n;

  func (p, hx, m, n, adata);
  nfev       = 1;
  p_eL2      = lm_l2nrmxmy (e, x, hx, n);
  init_p_eL2 = p_eL2;

  if (!isfinite (p_eL2))
    stop = 7;

  nu = 20; /* forces the first Jacobian */

  for (k = 0; k < itmax && !stop; k++)
  {
    if (p_eL2 <= eps3)
    {
      stop = 6;
      break;
    }

    if ((updp && nu > 16) || updjac == K)
    {
      if (using_ffdif)
      {
        /* misc_core.c:137-172 */
        for (j = 0; j < m; j++)
        {
          double d = 1e-04 * p[j];

          d = fabs (d);

          if (d < delta)
            d = delta;

          tmp   = p[j];
          p[j] += d;
          func (p, wrk, m, n, adata);
          p[j]  = tmp;
          d     = 1.0 / d;

          for (i = 0; i < n; i++)
            jac[i * m + j] = (wrk[i] - hx[i]) * d;
        }

        njap++;
        nfev += m;
      }
      else
      {
        /* misc_core.c:175-215 */
        for (j = 0; j < m; j++)
        {
          double d = 1e-04 * p[j];

          d = fabs (d);

          if (d < delta)
            d = delta;

          tmp   = p[j];
          p[j] -= d;
          func (p, wrk, m, n, adata);
          p[j]  = tmp + d;
          func (p, wrk2, m, n, adata);
          p[j]  = tmp;
          d     = 0.5 / d;

          for (i = 0; i < n; i++)
            jac[i * m + j] = (wrk2[i] - wrk[i]) * d;
        }

        njap++;
        nfev += 2 * m;
      }

      nu     = 2;
      updjac = 0;
      updp   = 0;
      newjac = 1;
    }

    if (newjac)
    {
      newjac = 0;

      for (i = 0; i < m * m; i++)
        jacTjac[i] = 0.0;

      for (i = 0; i < m; i++)
        jacTe[i] = 0.0;

      if (nm <= ORC_LM_BLOCKSZ_SQ)
      {
        for (l = n; l-- > 0;)
        {
          const double *jaclm = jac + l * m;

          for (i = m; i-- > 0;)
          {
            const double alpha = jaclm[i];

            for (j = i + 1; j-- > 0;)
              jacTjac[i * m + j] += jaclm[j] * alpha;

            jacTe[i] += alpha * e[l];
          }
        }
      }
      else
      {
        /* misc_core.c:82-129, the blocking multiply of the build without LAPACK: every entry is the sum, block of
         * 32 rows after block of 32 rows, of per-block partial sums (a build with LAPACK calls DGEMM here, whose
         * summation order belongs to the BLAS); then the upward J^T e loop of lm_core.c:651-660 */
        for (i = 0; i < m; i++)
          for (j = 0; j <= i; j++)
          {
            double b = 0.0;
            int kk;

            for (kk = 0; kk < n; kk += 32)
            {
              const int kend = (kk + 32 <= n) ? kk + 32 : n;
              double s = 0.0;

              for (l = kk; l < kend; l++)
                s += jac[l * m + j] * jac[l * m + i];

              b += s;
            }

            jacTjac[i * m + j] = b;
          }

        for (l = 0; l < n; l++)
        {
          const double el = e[l];

          for (i = 0; i < m; i++)
            jacTe[i] += jac[l * m + i] * el;
        }
      }

      for (i = m; i-- > 0;)
        for (j = i + 1; j < m; j++)
          jacTjac[i * m + j] = jacTjac[j * m + i];

      for (i = 0, p_L2 = jacTe_inf = 0.0; i < m; i++)
      {
        if (jacTe_inf < (tmp = fabs (jacTe[i])))
          jacTe_inf = tmp;

        diag[i] = jacTjac[i * m + i];
        p_L2   += p[i] * p[i];
      }
    }

    if (jacTe_inf <= eps1)
    {
      Dp_L2 = 0.0;
      stop  = 1;
      break;
    }

    if (k == 0)
    {
      for (i = 0, tmp = -DBL_MAX; i < m; i++)
        if (diag[i] > tmp)
          tmp = diag[i];

      mu = tau * tmp;
    }

    for (i = 0; i < m; i++)
      jacTjac[i * m + i] += mu;

    nlss++;

    if (lm_solve (jacTjac, jacTe, Dp, m))
    {
      for (i = 0, Dp_L2 = 0.0; i < m; i++)
      {
        pDp[i] = p[i] + (tmp = Dp[i]);
        Dp_L2 += tmp * tmp;
      }

      if (Dp_L2 <= eps2_sq * p_L2)
      {
        stop = 2;
        break;
      }

      if (Dp_L2 >= (p_L2 + eps2) / (1e-12 * 1e-12)) /* levmar's EPSILON = 1e-12 */
      {
        stop = 4;
        break;
      }

      func (pDp, wrk, m, n, adata);
      nfev++;
      pDp_eL2 = lm_l2nrmxmy (wrk2, x, wrk, n);

      if (!isfinite (pDp_eL2))
      {
        stop = 7;
        break;
      }

      dF = p_eL2 - pDp_eL2;

      if (updp || dF > 0)
      {
        /* Broyden rank-one update of the Jacobian */
        for (i = 0; i < n; i++)
        {
          for (l = 0, tmp = 0.0; l < m; l++)
            tmp += jac[i * m + l] * Dp[l];

          tmp = (wrk[i] - hx[i] - tmp) / Dp_L2;

          for (j = 0; j < m; j++)
            jac[i * m + j] += tmp * Dp[j];
        }

        updjac++;
        newjac = 1;
      }

      for (i = 0, dL = 0.0; i < m; i++)
        dL += Dp[i] * (mu * Dp[i] + jacTe[i]);

      if (dL > 0.0 && dF > 0.0)
      {
        tmp = (2.0 * dF / dL - 1.0);
        tmp = 1.0 - tmp * tmp * tmp;
        mu  = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334); /* levmar's ONE_THIRD */
        nu  = 2;

        for (i = 0; i < m; i++)
          p[i] = pDp[i];

        for (i = 0; i < n; i++)
        {
          e[i]  = wrk2[i];
          hx[i] = wrk[i];
        }

        p_eL2 = pDp_eL2;
        updp  = 1;
        continue;
      }
    }

    mu *= nu;
    nu2 = nu << 1;

    if (nu2 <= nu)
    {
      stop = 5;
      break;
    }

    nu = nu2;

    for (i = 0; i < m; i++)
      jacTjac[i * m + i] = diag[i];
  }

  if (k >= itmax)
    stop = 3;

  for (i = 0; i < m; i++)
    jacTjac[i * m + i] = diag[i];

  if (info != NULL)
  {
    info[0] = init_p_eL2;
    info[1] = p_eL2;
    info[2] = jacTe_inf;
    info[3] = Dp_L2;

    for (i = 0, tmp = -DBL_MAX; i < m; i++)
      if (tmp < jacTjac[i * m + i])
        tmp = jacTjac[i * m + i];

    info[4] = mu / tmp;
    info[5] = (double) k;
    info[6] = (double) stop;
    info[7] = (double) nfev;
    info[8] = (double) njap;
    info[9] = (double) nlss;
  }

  free (work);

  return (stop != 4 && stop != 7) ? k : -1;
}
