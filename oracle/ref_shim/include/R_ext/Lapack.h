/* ORACLE / TEST INFRASTRUCTURE ONLY: the four LAPACK routines the reference calls,
 * resolved against the OpenBLAS that ships inside scipy (LP64, scipy_ prefix). */
#ifndef SHIM_LAPACK_H
#define SHIM_LAPACK_H
#define F77_NAME(x) scipy_##x##_
#ifdef __cplusplus
extern "C" {
#endif
void scipy_dspev_(const char *jobz, const char *uplo, const int *n, double *ap, double *w,
                  double *z, const int *ldz, double *work, int *info);
void scipy_dspevx_(const char *jobz, const char *range, const char *uplo, const int *n, double *ap,
                   const double *vl, const double *vu, const int *il, const int *iu,
                   const double *abstol, int *m, double *w, double *z, const int *ldz, double *work,
                   int *iwork, int *ifail, int *info);
double scipy_dlamch_(const char *cmach);
void scipy_dgesvd_(const char *jobu, const char *jobvt, const int *m, const int *n, double *a,
                   const int *lda, double *s, double *u, const int *ldu, double *vt, const int *ldvt,
                   double *work, const int *lwork, int *info);
#ifdef __cplusplus
}
#endif
#endif
