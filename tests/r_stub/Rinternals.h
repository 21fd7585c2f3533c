/* tests/r_stub/Rinternals.h -- TEST INFRASTRUCTURE. A small stand-in for the part of R's C API that
 * raer_b200/csrc/r_shim.c uses, so that the R-facing shim is compiled and exercised in an image without R
 * (tests/test_r_shim.py). Names, argument order and semantics follow R's Rinternals.h; implementation in r_stub.c.
 * Never part of the product: where R is installed the shim is built against the real headers. */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
enum { NILSXP = 0, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22 };
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
SEXP Rf_allocVector(unsigned type, R_xlen_t n);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
double* REAL(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
int LENGTH(SEXP x);
R_xlen_t Rf_length(SEXP x);
SEXP Rf_mkChar(const char* s);
SEXP Rf_mkString(const char* s);
SEXP Rf_mkNamed(unsigned type, const char** names);
SEXP Rf_ScalarInteger(int v);
const char* Rf_translateChar(SEXP x);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val);
Rboolean Rf_isInteger(SEXP x);
Rboolean Rf_isReal(SEXP x);
Rboolean Rf_isLogical(SEXP x);
Rboolean Rf_isString(SEXP x);
Rboolean Rf_isNull(SEXP x);
Rboolean Rf_isMatrix(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
void Rf_error(const char* fmt, ...) __attribute__((noreturn));
void Rf_onintr(void) __attribute__((noreturn));
void REprintf(const char* fmt, ...);
char* R_alloc(size_t n, int size);
const char* R_ExpandFileName(const char* s);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
#define allocVector Rf_allocVector
#define length Rf_length
#define mkChar Rf_mkChar
#define mkString Rf_mkString
#define mkNamed Rf_mkNamed
#define ScalarInteger Rf_ScalarInteger
#define translateChar Rf_translateChar
#define setAttrib Rf_setAttrib
#define isInteger Rf_isInteger
#define isReal Rf_isReal
#define isLogical Rf_isLogical
#define isString Rf_isString
#define isNull Rf_isNull
#define isMatrix Rf_isMatrix
#define nrows Rf_nrows
#define ncols Rf_ncols
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#ifdef __cplusplus
}
#endif
#endif
