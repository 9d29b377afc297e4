/*
 * oracle/orc_blas.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Prototypes of the CBLAS / LAPACK entry points the reference path calls, as
 * exported (with a scipy_ prefix, LP64) by SciPy's bundled OpenBLAS
 * (scipy.libs/libscipy_openblas-*.so, OpenBLAS 0.3.31.dev).  The reference links
 * "whatever CBLAS/LAPACK meson finds" (meson.build:711-795); this is the one
 * present in this image.
 */
#ifndef ORC_BLAS_H
#define ORC_BLAS_H

enum { OrcRowMajor = 101, OrcColMajor = 102 };
enum { OrcNoTrans = 111, OrcTrans = 112 };
enum { OrcUpper = 121, OrcLower = 122 };
enum { OrcNonUnit = 131, OrcUnit = 132 };
enum { OrcLeft = 141, OrcRight = 142 };

void   scipy_cblas_daxpy (int n, double alpha, const double *x, int incx, double *y, int incy);
double scipy_cblas_ddot (int n, const double *x, int incx, const double *y, int incy);
double scipy_cblas_dnrm2 (int n, const double *x, int incx);
void   scipy_cblas_dscal (int n, double alpha, double *x, int incx);
void   scipy_cblas_dgemv (int order, int trans, int m, int n, double alpha, const double *a, int lda,
                          const double *x, int incx, double beta, double *y, int incy);
void   scipy_cblas_dsyrk (int order, int uplo, int trans, int n, int k, double alpha, const double *a, int lda,
                          double beta, double *c, int ldc);
void   scipy_cblas_dgemm (int order, int transa, int transb, int m, int n, int k, double alpha, const double *a, int lda,
                          const double *b, int ldb, double beta, double *c, int ldc);
void   scipy_cblas_dtrmv (int order, int uplo, int trans, int diag, int n, const double *a, int lda, double *x, int incx);
void   scipy_cblas_dtrsv (int order, int uplo, int trans, int diag, int n, const double *a, int lda, double *x, int incx);
void   scipy_cblas_dtrsm (int order, int side, int uplo, int trans, int diag, int m, int n, double alpha,
                          const double *a, int lda, double *b, int ldb);

void scipy_dpotrf_ (const char *uplo, const int *n, double *a, const int *lda, int *info);
void scipy_dposv_ (const char *uplo, const int *n, const int *nrhs, double *a, const int *lda, double *b, const int *ldb, int *info);
void scipy_dsysv_ (const char *uplo, const int *n, const int *nrhs, double *a, const int *lda, int *ipiv, double *b, const int *ldb,
                   double *work, const int *lwork, int *info);
void scipy_dgels_ (const char *trans, const int *m, const int *n, const int *nrhs, double *a, const int *lda, double *b, const int *ldb,
                   double *work, const int *lwork, int *info);
void scipy_dsyevr_ (const char *jobz, const char *range, const char *uplo, const int *n, double *a, const int *lda,
                    const double *vl, const double *vu, const int *il, const int *iu, const double *abstol, int *m, double *w,
                    double *z, const int *ldz, int *isuppz, double *work, const int *lwork, int *iwork, const int *liwork, int *info);
void scipy_openblas_set_num_threads (int n);
int  scipy_openblas_get_num_threads (void);

#endif
