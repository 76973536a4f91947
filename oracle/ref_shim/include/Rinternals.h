/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * A tiny in-process model of the handful of R C-API entry points the reference's
 * hot-path sources touch (there is no R in this image).  SEXPs are heap records
 * that are never collected (the driver process is short lived). */
#ifndef SHIM_RINTERNALS_H
#define SHIM_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
struct shim_sexp {
    int type;
    R_xlen_t len;
    void *data;
    struct shim_sexp *dim;
    struct shim_sexp *names;
};
typedef struct shim_sexp *SEXP;
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
extern double R_NaN, R_NaReal, R_PosInf, R_NegInf;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
#define NA_REAL R_NaReal
SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nr, int nc);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char *CHAR(SEXP);
SEXP Rf_mkChar(const char *);
SEXP Rf_mkString(const char *);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_isNull(SEXP);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarLogical(int);
SEXP Rf_duplicate(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
R_xlen_t XLENGTH(SEXP);
int LENGTH(SEXP);
int Rf_length(SEXP);
void Rprintf(const char *, ...);
void REprintf(const char *, ...);
void Rf_error(const char *, ...) __attribute__((noreturn));
void Rf_warning(const char *, ...);
int R_finite(double);
Rboolean R_ToplevelExec(void (*fun)(void *), void *data);
void R_CheckUserInterrupt(void);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#define R_FINITE(x) R_finite(x)
#define ISNA(x) ((x) != (x))
#define ISNAN(x) ((x) != (x))
/* the shim driver always passes vectors of the requested type already */
#define AS_INTEGER(x) (x)
#define AS_NUMERIC(x) (x)
#define NEW_NUMERIC(n) Rf_allocVector(REALSXP, n)
#define NEW_INTEGER(n) Rf_allocVector(INTSXP, n)
#define NEW_LOGICAL(n) Rf_allocVector(LGLSXP, n)
#define NEW_LIST(n) Rf_allocVector(VECSXP, n)
#define NEW_CHARACTER(n) Rf_allocVector(STRSXP, n)
#define SET_ELEMENT(x, i, v) SET_VECTOR_ELT(x, i, v)
#define GET_DIM(x) Rf_getAttrib(x, R_DimSymbol)
#define FCONE
#ifdef __cplusplus
}
#endif
#endif
