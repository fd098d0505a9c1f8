/* .Call shim between R and libgms_b200.so (include/gms_b200.h).
 *
 * Compiles only where R's headers exist (R CMD INSTALL; link with -lgms_b200). R is not installed in
 * the build image of this repository, so the file is syntax-checked there against the stub headers in
 * r-package/tests/rstub (tests/test_rglue_cpu.py) and EXECUTED on the GPU box against a minimal stand-in
 * for R's C API (r-package/tests/fakeR: the ~30 API functions this file uses, implemented in plain C; the
 * driver runs gmsr_create -> gmsr_set_* -> gmsr_predict / gmsr_tidy on the bundled data and
 * tests/test_gpu_rshim.py checks the vignette values). Under real R it has still never run.
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

SEXP gmsr_set_cache(SEXP ptr, SEXP enabled) {
    gms_handle* h = gmsr_get(ptr);
    gmsr_check(h, gms_set_cache(h, Rf_asInteger(enabled)));
    return R_NilValue;
}

static SEXP gmsr_array(int type, R_xlen_t n, int nd, const int* dims) {
    SEXP a = PROTECT(Rf_allocVector(type, n));
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, nd));
    for (int i = 0; i < nd; ++i) INTEGER(dim)[i] = dims[i];
    Rf_setAttrib(a, R_DimSymbol, dim);
    UNPROTECT(2);
    return a;
}

/* several effect sets (cross-validation folds) against ONE cross list: add/dom/asub are m x T x nsets arrays.
 * Returns list(out = double[ncomp, npairs, N, nsets], nseg = integer[N]). */
SEXP gmsr_predict_batched(SEXP ptr, SEXP model, SEXP nsets, SEXP ntraits, SEXP add, SEXP dom, SEXP asub, SEXP sire, SEXP dam) {
    gms_handle* h = gmsr_get(ptr);
    const int mdl = Rf_asInteger(model), S = Rf_asInteger(nsets), T = Rf_asInteger(ntraits);
    const R_xlen_t N = XLENGTH(sire);
    const int npairs = T * (T + 1) / 2, ncomp = mdl + 1;
    if (XLENGTH(dam) != N) Rf_error("gms_b200: sire and dam differ in length");
    const int dims[4] = {ncomp, npairs, (int)N, S};
    SEXP out = PROTECT(gmsr_array(REALSXP, (R_xlen_t)ncomp * npairs * N * S, 4, dims));
    SEXP nseg = PROTECT(Rf_allocVector(INTSXP, N));
    int rc = gms_predict_batched(h, mdl, S, T, REAL(add), Rf_isNull(dom) ? NULL : REAL(dom), Rf_isNull(asub) ? NULL : REAL(asub),
                                 (int64_t)N, INTEGER(sire), INTEGER(dam), REAL(out), INTEGER(nseg));
    if (rc != GMS_OK) { UNPROTECT(2); Rf_error("gms_b200: %s", gms_last_error(h)); }
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(res, 0, out);
    SET_VECTOR_ELT(res, 1, nseg);
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_STRING_ELT(nm, 0, Rf_mkChar("out"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("nseg"));
    Rf_setAttrib(res, R_NamesSymbol, nm);
    UNPROTECT(4);
    return res;
}

/* predictCrosses()' tidy table for the crosses of the last gmsr_predict on this handle (gms_tidy; R/gmsFunctions.R:726-871):
 * ntraits = T, siwts = double[T] or NULL, stdselint = numeric(1). Returns double[4, nslots, nof, N]:
 * (predMean, predVar, predSD, predUsefulness) x ([SELIND,] traits) x (BV [, TGV]) x cross. */
SEXP gmsr_tidy(SEXP ptr, SEXP model, SEXP ntraits, SEXP ncross, SEXP siwts, SEXP stdselint) {
    gms_handle* h = gmsr_get(ptr);
    const int mdl = Rf_asInteger(model), T = Rf_asInteger(ntraits), N = Rf_asInteger(ncross);
    const int nof = mdl == GMS_MODEL_A ? 1 : 2, nslots = T + (Rf_isNull(siwts) ? 0 : 1);
    if (!Rf_isNull(siwts) && LENGTH(siwts) != T) Rf_error("gms_b200: SIwts needs one weight per trait");
    const int dims[4] = {4, nslots, nof, N};
    SEXP out = PROTECT(gmsr_array(REALSXP, (R_xlen_t)4 * nslots * nof * N, 4, dims));
    int rc = gms_tidy(h, Rf_isNull(siwts) ? NULL : REAL(siwts), REAL(stdselint)[0], REAL(out));
    if (rc != GMS_OK) { UNPROTECT(1); Rf_error("gms_b200: %s", gms_last_error(h)); }
    UNPROTECT(1);
    return out;
}

/* selection-index variance w'Gw per cross and component from a gmsr_predict slab (gms_selind; R/gmsFunctions.R:825-853).
 * out: double[ncomp, npairs, N]; w: double[T]. Returns double[ncomp, N]. */
SEXP gmsr_selind(SEXP model, SEXP out, SEXP w) {
    const int mdl = Rf_asInteger(model), T = LENGTH(w), ncomp = mdl + 1, npairs = T * (T + 1) / 2;
    const R_xlen_t N = XLENGTH(out) / ((R_xlen_t)ncomp * npairs);
    const int dims[2] = {ncomp, (int)N};
    SEXP si = PROTECT(gmsr_array(REALSXP, (R_xlen_t)ncomp * N, 2, dims));
    if (gms_selind(mdl, T, (int64_t)N, REAL(out), REAL(w), REAL(si)) != GMS_OK) { UNPROTECT(1); Rf_error("gms_b200: gms_selind: bad arguments"); }
    UNPROTECT(1);
    return si;
}

/* predCrossMeans on the device-resident haplotypes (gms_cross_means; R/predCrossVar.R:336-399): add (and dom, or NULL):
 * m x T. Returns list(gebv = double[T, P], mean_bv = double[T, N], mean_tgv = double[T, N] or NULL). */
SEXP gmsr_cross_means(SEXP ptr, SEXP nparents, SEXP add, SEXP dom, SEXP sire, SEXP dam) {
    gms_handle* h = gmsr_get(ptr);
    const int T = Rf_ncols(add), P = Rf_asInteger(nparents);
    const R_xlen_t N = XLENGTH(sire);
    if (XLENGTH(dam) != N) Rf_error("gms_b200: sire and dam differ in length");
    const int dg[2] = {T, P}, dn[2] = {T, (int)N};
    SEXP gebv = PROTECT(gmsr_array(REALSXP, (R_xlen_t)T * P, 2, dg));
    SEXP mbv = PROTECT(gmsr_array(REALSXP, (R_xlen_t)T * N, 2, dn));
    SEXP mtgv = PROTECT(Rf_isNull(dom) ? R_NilValue : gmsr_array(REALSXP, (R_xlen_t)T * N, 2, dn));
    int rc = gms_cross_means(h, T, REAL(add), Rf_isNull(dom) ? NULL : REAL(dom), (int64_t)N, INTEGER(sire), INTEGER(dam),
                             REAL(gebv), REAL(mbv), Rf_isNull(dom) ? NULL : REAL(mtgv));
    if (rc != GMS_OK) { UNPROTECT(3); Rf_error("gms_b200: %s", gms_last_error(h)); }
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(res, 0, gebv);
    SET_VECTOR_ELT(res, 1, mbv);
    SET_VECTOR_ELT(res, 2, mtgv);
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_STRING_ELT(nm, 0, Rf_mkChar("gebv"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("mean_bv"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("mean_tgv"));
    Rf_setAttrib(res, R_NamesSymbol, nm);
    UNPROTECT(5);
    return res;
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
    {"gmsr_set_cache", (DL_FUNC)&gmsr_set_cache, 2},
    {"gmsr_predict_batched", (DL_FUNC)&gmsr_predict_batched, 9},
    {"gmsr_tidy", (DL_FUNC)&gmsr_tidy, 6},
    {"gmsr_selind", (DL_FUNC)&gmsr_selind, 3},
    {"gmsr_cross_means", (DL_FUNC)&gmsr_cross_means, 6},
    {NULL, NULL, 0}};

void R_init_genomicMateSelectRb200(DllInfo* dll) {
    R_registerRoutines(dll, NULL, gmsr_calls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
