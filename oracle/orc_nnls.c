/*
 * oracle/orc_nnls.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of ncm_nnls_solve with NCM_NNLS_UMETHOD_NORMAL
 * (numcosmo/ncm/algebra/ncm_nnls.c:544-568, 573-638, 655-666, 710-751, 767-871)
 * and of the NcmISet operations it uses (numcosmo/ncm/core/ncm_iset.c:376-404,
 * 463-512, 567-625, 853-884, 920-985, 994-1077, 1180-1204).  The index set is kept
 * as an ascending int array: every reference operation sorts the GQueue before
 * reading it (_ncm_iset_sort), so the observable order is the same.
 *
 * gsl_sort_vector_smallest_index / gsl_sort_largest_index (GSL sort/subsetind_source.c,
 * not vendored) are restated from the published algorithm (insertion into a
 * bounded sorted list, first-come wins on ties).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ncm_oracle.h"
#include "orc_blas.h"

/* GSL sort/subsetind_source.c: gsl_sort_smallest_index */
void
orc_sort_smallest_index (int *p, int k, const double *src, int stride, int n)
{
  int i, j;
  double xbound;

  if ((k == 0) || (n == 0))
    return;

  j      = 1;
  xbound = src[0 * stride];
  p[0]   = 0;

  for (i = 1; i < n; i++)
  {
    int i1;
    double xi = src[i * stride];

    if (j < k)
      j++;
    else if (xi >= xbound)
      continue;

    for (i1 = j - 1; i1 > 0; i1--)
    {
      if (xi > src[stride * p[i1 - 1]])
        break;

      p[i1] = p[i1 - 1];
    }

    p[i1] = i;

    xbound = src[stride * p[j - 1]];
  }
}

/* GSL sort/subsetind_source.c: gsl_sort_largest_index */
void
orc_sort_largest_index (int *p, int k, const double *src, int stride, int n)
{
  int i, j;
  double xbound;

  if ((k == 0) || (n == 0))
    return;

  j      = 1;
  xbound = src[0 * stride];
  p[0]   = 0;

  for (i = 1; i < n; i++)
  {
    int i1;
    double xi = src[i * stride];

    if (j < k)
      j++;
    else if (xi <= xbound)
      continue;

    for (i1 = j - 1; i1 > 0; i1--)
    {
      if (xi < src[stride * p[i1 - 1]])
        break;

      p[i1] = p[i1 - 1];
    }

    p[i1] = i;

    xbound = src[stride * p[j - 1]];
  }
}

typedef struct orc_iset
{
  int n;     /* max size */
  int len;
  int *idx;  /* ascending */
} orc_iset;

static void
iset_init (orc_iset *s, int n)
{
  s->n   = n;
  s->len = 0;
  s->idx = (int *) malloc (sizeof (int) * (n > 0 ? n : 1));
}

static void
iset_copy (const orc_iset *s, orc_iset *t)
{
  t->len = s->len;
  memcpy (t->idx, s->idx, sizeof (int) * s->len);
}

static int
cmp_int (const void *a, const void *b)
{
  const int ia = *(const int *) a, ib = *(const int *) b;

  return (ia > ib) - (ia < ib);
}

static void
iset_del (orc_iset *s, int i)
{
  int j;

  for (j = 0; j < s->len; j++)
  {
    if (s->idx[j] == i)
    {
      memmove (&s->idx[j], &s->idx[j + 1], sizeof (int) * (s->len - j - 1));
      s->len--;

      return;
    }
  }
}

typedef struct orc_nnls
{
  int nrows, ncols, lda;
  const double *A;
  double *M, *M_U, *b, *x_tmp, *x_try, *residuals, *residuals_try, *mgrad, *tmp;
  int *atmp, *ptmp;
  orc_iset Pset, Pset_try, invalid;
  int uncols;
  orc_nnls_stats *st;
} orc_nnls;

/* ncm_nnls.c:544-568: gather M[P,P], b[P] (ncm_iset.c:463-512, 567-625); sub views keep tda = ncols */
static void
prepare_usys_normal (orc_nnls *self, const orc_iset *Pset)
{
  const int n = self->ncols;

  self->uncols = Pset->len;

  if (self->uncols == n)
  {
    memcpy (self->M_U, self->M, sizeof (double) * n * n);
    memcpy (self->x_tmp, self->b, sizeof (double) * n);
  }
  else
  {
    int i, j;

    for (i = 0; i < Pset->len; i++)
    {
      const int k = Pset->idx[i];

      for (j = 0; j < Pset->len; j++)
      {
        const int l = Pset->idx[j];

        self->M_U[i * n + j] = self->M[k * n + l];
      }

      self->x_tmp[i] = self->b[k];
    }
  }
}

/* ncm_nnls.c:608-638 via dgels on A[:,P]; equivalent restatement (column-major pack, trans = 'N') */
static void
solve_normal_QR (orc_nnls *self, const orc_iset *Pset, const double *f)
{
  const int m = self->nrows, n = Pset->len;
  const int ldb = (m > n) ? m : n;
  double *Acm = (double *) malloc (sizeof (double) * m * n);
  double *rhs = (double *) calloc (ldb, sizeof (double));
  double wq;
  int lwork = -1, info = 0, nrhs = 1, i, j;
  double *work;

  self->uncols = n;

  for (j = 0; j < n; j++)
    for (i = 0; i < m; i++)
      Acm[j * m + i] = self->A[i * self->lda + Pset->idx[j]];

  memcpy (rhs, f, sizeof (double) * m);

  scipy_dgels_ ("N", &m, &n, &nrhs, Acm, &m, rhs, &ldb, &wq, &lwork, &info);
  lwork = (int) wq;
  work  = (double *) malloc (sizeof (double) * lwork);
  scipy_dgels_ ("N", &m, &n, &nrhs, Acm, &m, rhs, &ldb, work, &lwork, &info);

  memcpy (self->x_tmp, rhs, sizeof (double) * n);

  free (work);
  free (rhs);
  free (Acm);

  if (self->st != NULL)
    self->st->n_qr++;
}

/* ncm_nnls.c:573-606 */
static void
solve_normal_LU (orc_nnls *self, const orc_iset *Pset, const double *f)
{
  int info = 0, nrhs = 1, lwork = -1;
  const int lda = self->ncols;
  int *ipiv;
  double wq, *work;

  prepare_usys_normal (self, Pset);

  ipiv = (int *) malloc (sizeof (int) * self->ncols);

  /* 'U' row-major == 'L' column-major (_NCM_LAPACK_CONV_UPLO, ncm_lapack.c:58,798) */
  scipy_dsysv_ ("L", &self->uncols, &nrhs, self->M_U, &lda, ipiv, self->x_tmp, &self->uncols, &wq, &lwork, &info);
  lwork = (int) wq;
  work  = (double *) malloc (sizeof (double) * lwork);
  scipy_dsysv_ ("L", &self->uncols, &nrhs, self->M_U, &lda, ipiv, self->x_tmp, &self->uncols, work, &lwork, &info);

  free (work);
  free (ipiv);

  if (self->st != NULL)
    self->st->n_lu++;

  if (info > 0)
    solve_normal_QR (self, Pset, f);
}

/* ncm_nnls.c:655-666 ; ncm_matrix.c:1199-1210 (dposv) */
static void
solve_normal_cholesky (orc_nnls *self, const orc_iset *Pset, const double *f)
{
  int info = 0, nrhs = 1;
  const int lda = self->ncols;

  prepare_usys_normal (self, Pset);

  scipy_dposv_ ("L", &self->uncols, &nrhs, self->M_U, &lda, self->x_tmp, &self->uncols, &info);

  if (self->st != NULL)
    self->st->n_chol++;

  if (getenv ("ORC_NNLS_TRACE") != NULL)   /* debugging aid: passive-set size of every factorisation */
    fprintf (stderr, "orc_nnls: chol |P| = %d info = %d\n", self->uncols, info);

  if (getenv ("ORC_NNLS_DUMP") != NULL)    /* debugging aid: append (|P|, indices, solution) of every factorisation to a binary file */
  {
    FILE *fp = fopen (getenv ("ORC_NNLS_DUMP"), "ab");

    if (fp != NULL)
    {
      fwrite (&Pset->len, sizeof (int), 1, fp);
      fwrite (Pset->idx, sizeof (int), Pset->len, fp);
      fwrite (self->x_tmp, sizeof (double), Pset->len, fp);
      fclose (fp);
    }
  }

  if (info > 0)
    solve_normal_LU (self, Pset, f);
}

/* ncm_nnls.c:710-720 */
static double
compute_residuals (orc_nnls *self, const double *x, const double *f, double *residuals)
{
  memcpy (residuals, f, sizeof (double) * self->nrows);
  scipy_cblas_dgemv (OrcRowMajor, OrcNoTrans, self->nrows, self->ncols, -1.0, self->A, self->lda, x, 1, 1.0, residuals, 1);

  return scipy_cblas_dnrm2 (self->nrows, residuals, 1);
}

/* ncm_nnls.c:722-726 */
static void
compute_mgrad (orc_nnls *self, const double *residuals, double *mgrad)
{
  scipy_cblas_dgemv (OrcRowMajor, OrcTrans, self->nrows, self->ncols, 1.0, self->A, self->lda, residuals, 1, 0.0, mgrad, 1);
}

/* ncm_iset.c:853-884 */
static void
iset_get_subset_vec_lt (const orc_iset *s, orc_iset *out, const double *v, const double tol)
{
  int j;

  out->len = 0;

  for (j = 0; j < s->len; j++)
  {
    const int i = s->idx[j];

    if (v[i] < tol)
      out->idx[out->len++] = i;
  }
}

/* ncm_iset.c:920-985 */
static int
iset_remove_smallest_subset (orc_nnls *self, const orc_iset *invalid, orc_iset *target, const double *v, int max_remove)
{
  const int rsize = invalid->len;
  int j;

  if (max_remove >= rsize)
  {
    for (j = 0; j < rsize; j++)
      iset_del (target, invalid->idx[j]);

    return rsize;
  }
  else
  {
    for (j = 0; j < rsize; j++)
    {
      self->tmp[j]  = v[invalid->idx[j]];
      self->atmp[j] = invalid->idx[j];
    }

    orc_sort_smallest_index (self->ptmp, max_remove, self->tmp, 1, rsize);

    for (j = 0; j < max_remove; j++)
    {
      const int vi = self->ptmp[j];
      const int ti = self->atmp[vi];

      iset_del (target, ti);
    }

    return max_remove;
  }
}

/* ncm_iset.c:994-1077 */
static int
iset_add_largest_subset (orc_nnls *self, orc_iset *s, const double *v, const double min, const double add_frac)
{
  const int max_size = s->n;
  const int rsize    = s->len;
  const int csize    = max_size - rsize;
  int adds, j, k, node;

  if (csize == 0)
    return 0;

  j = 0;
  k = 0;

  for (node = 0; node < s->len; node++)
  {
    const int i = s->idx[node];

    for ( ; j < i; j++)
    {
      const double v_j = v[j];

      if (v_j > min)
      {
        self->tmp[k]  = v_j;
        self->atmp[k] = j;
        k++;
      }
    }

    j = i + 1;
  }

  for ( ; j < max_size; j++)
  {
    const double v_j = v[j];

    if (v_j > min)
    {
      self->tmp[k]  = v_j;
      self->atmp[k] = j;
      k++;
    }
  }

  {
    /* adds = MIN (k, MAX (k * add_frac, 1)) evaluated in double, truncated to guint */
    double a = k * add_frac;

    if (a < 1.0)
      a = 1.0;

    if ((double) k < a)
      a = (double) k;

    adds = (int) a;
  }

  if (adds > 0)
  {
    orc_sort_largest_index (self->ptmp, adds, self->tmp, 1, k);

    for (j = 0; j < adds; j++)
    {
      const int vi = self->ptmp[j];
      const int ti = self->atmp[vi];

      s->idx[s->len++] = ti;
    }

    qsort (s->idx, s->len, sizeof (int), cmp_int);
  }

  return adds;
}

/* ncm_nnls.c:728-751 */
static void
solve_feasible (orc_nnls *self, orc_iset *Pset, double *x, const double *f, int max_remove)
{
  int j;

  solve_normal_cholesky (self, Pset, f);

  memset (x, 0, sizeof (double) * self->ncols);

  for (j = 0; j < Pset->len; j++)
    x[Pset->idx[j]] = self->x_tmp[j];

  iset_get_subset_vec_lt (Pset, &self->invalid, x, 1.0e-300);

  while (self->invalid.len)
  {
    iset_remove_smallest_subset (self, &self->invalid, Pset, x, max_remove);

    solve_normal_cholesky (self, Pset, f);

    memset (x, 0, sizeof (double) * self->ncols);

    for (j = 0; j < Pset->len; j++)
      x[Pset->idx[j]] = self->x_tmp[j];

    iset_get_subset_vec_lt (Pset, &self->invalid, x, 1.0e-300);
  }
}

/* ncm_nnls.c:767-871 */
double
orc_nnls_solve (const double *A, int nrows, int ncols, int lda, double *x, const double *f, double reltol, orc_nnls_stats *stats)
{
  orc_nnls S, *self = &S;
  const size_t nn   = (size_t) ncols * ncols;
  const int big     = (nrows > ncols) ? nrows : ncols;
  double rnorm;
  int i;

  memset (self, 0, sizeof (S));
  self->nrows = nrows;
  self->ncols = ncols;
  self->lda   = lda;
  self->A     = A;
  self->st    = stats;

  if (stats != NULL)
    memset (stats, 0, sizeof (*stats));

  self->M             = (double *) malloc (sizeof (double) * nn);
  self->M_U           = (double *) malloc (sizeof (double) * nn);
  self->b             = (double *) malloc (sizeof (double) * big);
  self->x_tmp         = (double *) malloc (sizeof (double) * big);
  self->x_try         = (double *) malloc (sizeof (double) * ncols);
  self->residuals     = (double *) malloc (sizeof (double) * big);
  self->residuals_try = (double *) malloc (sizeof (double) * big);
  self->mgrad         = (double *) malloc (sizeof (double) * ncols);
  self->tmp           = (double *) malloc (sizeof (double) * ncols);
  self->atmp          = (int *) malloc (sizeof (int) * ncols);
  self->ptmp          = (int *) malloc (sizeof (int) * ncols);
  iset_init (&self->Pset, ncols);
  iset_init (&self->Pset_try, ncols);
  iset_init (&self->invalid, ncols);
  memset (self->M, 0, sizeof (double) * nn);

  /* ncm_matrix_square_to_sym (A, 'T', 'U', M) ; ncm_matrix_update_vector (A, 'T', 1.0, f, 0.0, b) */
  scipy_cblas_dsyrk (OrcRowMajor, OrcUpper, OrcTrans, ncols, nrows, 1.0, A, lda, 0.0, self->M, ncols);
  scipy_cblas_dgemv (OrcRowMajor, OrcTrans, nrows, ncols, 1.0, A, lda, f, 1, 0.0, self->b, 1);

  self->Pset.len = 0;

  for (i = 0; i < ncols; i++)
    self->Pset.idx[self->Pset.len++] = i;

  solve_feasible (self, &self->Pset, x, f, ncols);
  rnorm = compute_residuals (self, x, f, self->residuals);
  compute_mgrad (self, self->residuals, self->mgrad);

  while (1)
  {
    double add_frac = 1.0;
    int finish      = 0;
    double lrnorm;
    int added;

    while (1)
    {
      iset_copy (&self->Pset, &self->Pset_try);
      added = iset_add_largest_subset (self, &self->Pset_try, self->mgrad, rnorm * reltol, add_frac);

      add_frac *= 0.5;

      if (added == 0)
      {
        finish = 1;
        break;
      }

      solve_feasible (self, &self->Pset_try, self->x_try, f, added);
      lrnorm = compute_residuals (self, self->x_try, f, self->residuals_try);

      if (rnorm - lrnorm > rnorm * reltol)
      {
        iset_copy (&self->Pset_try, &self->Pset);
        memcpy (x, self->x_try, sizeof (double) * ncols);
        memcpy (self->residuals, self->residuals_try, sizeof (double) * nrows);
        rnorm = lrnorm;

        if (stats != NULL)
          stats->n_outer++;

        break;
      }

      if (added == 1)
      {
        finish = 1;
        break;
      }
    }

    if (finish)
      break;

    compute_mgrad (self, self->residuals, self->mgrad);
  }

  if (stats != NULL)
    stats->n_passive = self->Pset.len;

  free (self->M);
  free (self->M_U);
  free (self->b);
  free (self->x_tmp);
  free (self->x_try);
  free (self->residuals);
  free (self->residuals_try);
  free (self->mgrad);
  free (self->tmp);
  free (self->atmp);
  free (self->ptmp);
  free (self->Pset.idx);
  free (self->Pset_try.idx);
  free (self->invalid.idx);

  return rnorm;
}
