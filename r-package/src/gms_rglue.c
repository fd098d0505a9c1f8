/* .Call shim between R and libgms_b200.so (include/gms_b200.h).
 *
 * Compiles only where R's headers exist (R CMD INSTALL; link with -lgms_b200). R is not installed in
 * the build image of this repository, so the file is syntax-checked there against the stub headers in
 * r-package/tests/rstub (tests/test_rglue_cpu.py) and exercised for real only by a user with R.
 *
 * Conventions: every entry point takes/returns SEXPs, touches R's API only on R's main thread,
 * unpacks INTEGER()/REAL() pointers without copying (R matrices are column-major, which is what the
 * C ABI takes), allocates results with allocVector under PROTECT, and turns a non-zero status into an
 * R error carrying gms_last_error(). The engine handle lives in an external pointer with a finalizer,
 * so the HBM-resident tiles and haplotypes persist across the many predCrossVars() calls of a
 * cross-validation run (R/parentWiseCrossVal.R:430-490). */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>

#include "gms_b200.h"

static void gmsr_finalize(SEXP ptr) {
    gms_handle* h = (gms_handle*)R_ExternalPtrAddr(ptr);
    if (h) { gms_destroy(h); R_ClearExternalPtr(ptr); }
}

static gms_handle* gmsr_get(SEXP ptr) {
    gms_handle* h = (gms_handle*)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("gms_b200: engine handle is closed");
    return h;
}

static void gmsr_check(gms_handle* h, int rc) {
    if (rc != GMS_OK) Rf_error("gms_b200: %s", gms_last_error(h));
}

SEXP gmsr_create(SEXP device) {
    gms_handle* h = NULL;
    int rc = gms_create(&h, Rf_asInteger(device));
    if (rc != GMS_OK) Rf_error("gms_b200: %s", gms_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, gmsr_finalize, TRUE);
    UNPROTECT(1);
    return ptr;
}

SEXP gmsr_close(SEXP ptr) { gmsr_finalize(ptr); return R_NilValue; }

/* m_c: integer vector (SNPs per chromosome); cM: double vector, chromosome-major */
SEXP gmsr_set_genmap(SEXP ptr, SEXP m_c, SEXP cM) {
    gms_handle* h = gmsr_get(ptr);
    int C = LENGTH(m_c);
    int64_t* mc = (int64_t*)R_alloc(C, sizeof(int64_t));
    for (int c = 0; c < C; ++c) mc[c] = INTEGER(m_c)[c];
    gmsr_check(h, gms_set_genmap(h, C, mc, REAL(cM)));
    return R_NilValue;
}

/* blocks: list of square double matrices (the diagonal blocks of recombFreqMat = 1-2c) */
SEXP gmsr_set_recomb_blocks(SEXP ptr, SEXP blocks) {
    gms_handle* h = gmsr_get(ptr);
    int C = LENGTH(blocks);
    int64_t* mc = (int64_t*)R_alloc(C, sizeof(int64_t));
    const double** bp = (const double**)R_alloc(C, sizeof(double*));
    for (int c = 0; c < C; ++c) {
        SEXP b = VECTOR_ELT(blocks, c);
        if (!Rf_isReal(b) || !Rf_isMatrix(b) || Rf_nrows(b) != Rf_ncols(b)) Rf_error("gms_b200: blocks must be square double matrices");
        mc[c] = Rf_nrows(b);
        bp[c] = REAL(b);
    }
    gmsr_check(h, gms_set_recomb_blocks(h, C, mc, bp));
    return R_NilValue;
}

/* hap: integer matrix (2P x m), rows ordered HapA, HapB per parent (the glue reorders by name) */
SEXP gmsr_set_haplotypes(SEXP ptr, SEXP hap) {
    gms_handle* h = gmsr_get(ptr);
    if (!Rf_isInteger(hap) || !Rf_isMatrix(hap) || Rf_nrows(hap) % 2) Rf_error("gms_b200: haploMat must be an integer matrix with 2 rows per parent");
    gmsr_check(h, gms_set_haplotypes_i32cm(h, Rf_nrows(hap) / 2, Rf_ncols(hap), INTEGER(hap)));
    return R_NilValue;
}

/* model: 0 A, 1 AD, 2 DirDom; add/dom/asub: m x T double matrices (dom/asub may be NULL);
 * sire/dam: 0-based integer parent indices. Returns list(out = double[ncomp, npairs, N], nseg = integer[N])
 * (out is filled cross-major by the C ABI, i.e. an R array with dim c(ncomp, npairs, N)). */
SEXP gmsr_predict(SEXP ptr, SEXP model, SEXP add, SEXP dom, SEXP asub, SEXP sire, SEXP dam) {
    gms_handle* h = gmsr_get(ptr);
    const int mdl = Rf_asInteger(model), T = Rf_ncols(add);
    const R_xlen_t N = XLENGTH(sire);
    const int npairs = T * (T + 1) / 2, ncomp = mdl + 1;
    if (XLENGTH(dam) != N) Rf_error("gms_b200: sire and dam differ in length");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)ncomp * npairs * N));
    SEXP nseg = PROTECT(Rf_allocVector(INTSXP, N));
    int rc = gms_predict(h, mdl, T, REAL(add), Rf_isNull(dom) ? NULL : REAL(dom), Rf_isNull(asub) ? NULL : REAL(asub),
                         (int64_t)N, INTEGER(sire), INTEGER(dam), REAL(out), INTEGER(nseg));
    if (rc != GMS_OK) { UNPROTECT(2); Rf_error("gms_b200: %s", gms_last_error(h)); }
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, 3));
    INTEGER(dim)[0] = ncomp; INTEGER(dim)[1] = npairs; INTEGER(dim)[2] = (int)N;
    Rf_setAttrib(out, R_DimSymbol, dim);
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(res, 0, out);
    SET_VECTOR_ELT(res, 1, nseg);
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_STRING_ELT(nm, 0, Rf_mkChar("out"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("nseg"));
    Rf_setAttrib(res, R_NamesSymbol, nm);
    UNPROTECT(5);
    return res;
}

/* same as gmsr_predict on a LIST of handles (one per GPU, same map + haplotypes): gms_predict_sharded cuts
 * the cross list into contiguous shards, one host thread per GPU, all joined before returning to R */
SEXP gmsr_predict_sharded(SEXP ptrs, SEXP model, SEXP add, SEXP dom, SEXP asub, SEXP sire, SEXP dam) {
    const int nh = LENGTH(ptrs);
    if (nh < 1) Rf_error("gms_b200: no engine handles");
    gms_handle** hs = (gms_handle**)R_alloc(nh, sizeof(gms_handle*));
    for (int i = 0; i < nh; ++i) hs[i] = gmsr_get(VECTOR_ELT(ptrs, i));
    const int mdl = Rf_asInteger(model), T = Rf_ncols(add);
    const R_xlen_t N = XLENGTH(sire);
    const int npairs = T * (T + 1) / 2, ncomp = mdl + 1;
    if (XLENGTH(dam) != N) Rf_error("gms_b200: sire and dam differ in length");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)ncomp * npairs * N));
    SEXP nseg = PROTECT(Rf_allocVector(INTSXP, N));
    int rc = gms_predict_sharded(hs, nh, mdl, T, REAL(add), Rf_isNull(dom) ? NULL : REAL(dom),
                                 Rf_isNull(asub) ? NULL : REAL(asub), (int64_t)N, INTEGER(sire), INTEGER(dam),
                                 REAL(out), INTEGER(nseg));
    if (rc != GMS_OK) { UNPROTECT(2); Rf_error("gms_b200: %s", gms_last_error(hs[0])); }
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, 3));
    INTEGER(dim)[0] = ncomp; INTEGER(dim)[1] = npairs; INTEGER(dim)[2] = (int)N;
    Rf_setAttrib(out, R_DimSymbol, dim);
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(res, 0, out);
    SET_VECTOR_ELT(res, 1, nseg);
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_STRING_ELT(nm, 0, Rf_mkChar("out"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("nseg"));
    Rf_setAttrib(res, R_NamesSymbol, nm);
    UNPROTECT(5);
    return res;
}

SEXP gmsr_set_mode(SEXP ptr, SEXP mode) {
    gms_handle* h = gmsr_get(ptr);
    gmsr_check(h, gms_set_mode(h, Rf_asInteger(mode)));
    return R_NilValue;
}

static const R_CallMethodDef gmsr_calls[] = {
    {"gmsr_create", (DL_FUNC)&gmsr_create, 1},
    {"gmsr_close", (DL_FUNC)&gmsr_close, 1},
    {"gmsr_set_genmap", (DL_FUNC)&gmsr_set_genmap, 3},
    {"gmsr_set_recomb_blocks", (DL_FUNC)&gmsr_set_recomb_blocks, 2},
    {"gmsr_set_haplotypes", (DL_FUNC)&gmsr_set_haplotypes, 2},
    {"gmsr_predict", (DL_FUNC)&gmsr_predict, 7},
    {"gmsr_predict_sharded", (DL_FUNC)&gmsr_predict_sharded, 7},
    {"gmsr_set_mode", (DL_FUNC)&gmsr_set_mode, 2},
    {NULL, NULL, 0}};

void R_init_genomicMateSelectRb200(DllInfo* dll) {
    R_registerRoutines(dll, NULL, gmsr_calls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
