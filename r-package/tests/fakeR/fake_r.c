/* A minimal stand-in for R's C API: just enough object model (typed vectors with attributes, external pointers,
 * a protect stack that only counts) to EXECUTE r-package/src/gms_rglue.c without R. Test infrastructure only
 * (tests/test_gpu_rshim.py); declarations are those of r-package/tests/rstub/Rinternals.h, semantics follow
 * "Writing R Extensions" for the subset the shim uses. Rf_error prints the message and exits with status 3,
 * which is how the driver's caller sees an R condition. */
#define _POSIX_C_SOURCE 200809L
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "R_ext/Rdynload.h"

#define CHARSXP 9
#define EXTPTRSXP 22
#define NILSXP 0

struct SEXPREC {
    int type;
    R_xlen_t length;
    void* data;          /* INTSXP: int[], REALSXP: double[], VECSXP / STRSXP: SEXP[], CHARSXP: char[], EXTPTRSXP: address */
    SEXP dim, names;     /* the two attributes the shim sets / reads */
    R_CFinalizer_t fin;
};

static struct SEXPREC nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL};
static struct SEXPREC dim_sym = {CHARSXP, 3, "dim", NULL, NULL, NULL};
static struct SEXPREC names_sym = {CHARSXP, 5, "names", NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj, R_DimSymbol = &dim_sym, R_NamesSymbol = &names_sym;
static int protect_depth = 0;

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    size_t esz = type == INTSXP ? sizeof(int) : type == REALSXP ? sizeof(double) : sizeof(SEXP);
    s->type = (int)type;
    s->length = n;
    s->data = calloc((size_t)(n > 0 ? n : 1), esz);
    s->dim = s->names = R_NilValue;
    if (type == VECSXP || type == STRSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_mkChar(const char* str) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = CHARSXP;
    s->length = (R_xlen_t)strlen(str);
    s->data = strdup(str);
    s->dim = s->names = R_NilValue;
    return s;
}
int* INTEGER(SEXP s) { if (s->type != INTSXP) Rf_error("INTEGER() on a non-integer object"); return (int*)s->data; }
double* REAL(SEXP s) { if (s->type != REALSXP) Rf_error("REAL() on a non-double object"); return (double*)s->data; }
int LENGTH(SEXP s) { return (int)s->length; }
R_xlen_t XLENGTH(SEXP s) { return s->length; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { if (s->type != VECSXP || i >= s->length) Rf_error("VECTOR_ELT out of range"); return ((SEXP*)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { if (s->type != VECSXP || i >= s->length) Rf_error("SET_VECTOR_ELT out of range"); ((SEXP*)s->data)[i] = v; return v; }
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) { if (s->type != STRSXP || i >= s->length) Rf_error("SET_STRING_ELT out of range"); ((SEXP*)s->data)[i] = v; }
SEXP Rf_setAttrib(SEXP s, SEXP name, SEXP v) {
    if (name == R_DimSymbol) s->dim = v; else if (name == R_NamesSymbol) s->names = v; else Rf_error("unsupported attribute");
    return v;
}
SEXP Rf_protect(SEXP s) { ++protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; if (protect_depth < 0) Rf_error("protect stack underflow"); }
int fakeR_protect_depth(void) { return protect_depth; }
int Rf_asInteger(SEXP s) { return s->type == INTSXP ? ((int*)s->data)[0] : (int)((double*)s->data)[0]; }
Rboolean Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
Rboolean Rf_isReal(SEXP s) { return s->type == REALSXP; }
Rboolean Rf_isInteger(SEXP s) { return s->type == INTSXP; }
Rboolean Rf_isMatrix(SEXP s) { return s->dim != R_NilValue && s->dim->length == 2; }
int Rf_nrows(SEXP s) { return Rf_isMatrix(s) ? ((int*)s->dim->data)[0] : (int)s->length; }
int Rf_ncols(SEXP s) { return Rf_isMatrix(s) ? ((int*)s->dim->data)[1] : 1; }
void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Error: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(3);
}
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }     /* R frees these at the end of .Call; the driver is short-lived */
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = EXTPTRSXP;
    s->data = p;
    s->dim = s->names = R_NilValue;
    return s;
}
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fin, Rboolean onexit) { (void)onexit; s->fin = fin; }
void fakeR_run_finalizer(SEXP s) { if (s->type == EXTPTRSXP && s->fin) s->fin(s); }

/* .Call dispatch through the table the shim registers in R_init_<pkg> */
static const R_CallMethodDef* call_table = NULL;
int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
    (void)dll; (void)c; (void)f; (void)e;
    call_table = call;
    return 1;
}
int R_useDynamicSymbols(DllInfo* dll, int v) { (void)dll; (void)v; return 0; }
DL_FUNC fakeR_lookup(const char* name, int nargs) {
    for (const R_CallMethodDef* d = call_table; d && d->name; ++d)
        if (!strcmp(d->name, name)) {
            if (d->numArgs != nargs) Rf_error("%s registered with %d arguments, called with %d", name, d->numArgs, nargs);
            return d->fun;
        }
    Rf_error("%s is not a registered .Call routine", name);
}
SEXP fakeR_names(SEXP s) { return s->names; }
SEXP fakeR_dim(SEXP s) { return s->dim; }
const char* fakeR_string_elt(SEXP s, int i) { return (s->type == STRSXP && i < s->length) ? (const char*)((SEXP*)s->data)[i]->data : ""; }
