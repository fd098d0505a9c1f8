/* Minimal stand-in for R's C API, ONLY so that r-package/src/gms_rglue.c can be syntax- and
 * type-checked in an image without R (tests/test_rglue_cpu.py: gcc -fsyntax-only). Declarations
 * follow R's public headers; nothing here is linked or executed. */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
int* INTEGER(SEXP); double* REAL(SEXP);
int LENGTH(SEXP); R_xlen_t XLENGTH(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t); SEXP Rf_mkChar(const char*); SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
int Rf_asInteger(SEXP); int Rf_nrows(SEXP); int Rf_ncols(SEXP);
Rboolean Rf_isNull(SEXP); Rboolean Rf_isReal(SEXP); Rboolean Rf_isInteger(SEXP); Rboolean Rf_isMatrix(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
char* R_alloc(size_t, int);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP); void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
