/* Driver: plays the part of r-package/R/predCrossVars.R + predictCrosses.R for one call, through the REGISTERED
 * .Call entry points of r-package/src/gms_rglue.c, on the fake R API of fake_r.c. Test infrastructure only.
 *
 *   run_shim <problem.bin> <result.bin>
 * problem.bin (little endian, written by tests/test_gpu_rshim.py):
 *   int32 model, C, m, P, T, N, use_map, selind; double stdselint; int32 m_c[C]; then
 *   use_map ? double cM[m] : double blocks (C column-major m_c x m_c matrices);
 *   int32 hap[2P x m] column-major; double add[m x T]; dom[m x T] if model >= 1; asub[m x T] if model == 2;
 *   int32 sire[N], dam[N]; double siwts[T] if selind.
 * result.bin: int32 ncomp, npairs, N, nof, nslots; double out[ncomp x npairs x N]; int32 nseg[N]; double tidy[4 x nslots x nof x N];
 *   int32 protect_depth (must be 0). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "R_ext/Rdynload.h"

void R_init_genomicMateSelectRb200(DllInfo*);
DL_FUNC fakeR_lookup(const char*, int);
int fakeR_protect_depth(void);
void fakeR_run_finalizer(SEXP);
SEXP fakeR_names(SEXP);
const char* fakeR_string_elt(SEXP, int);

typedef SEXP (*F1)(SEXP);
typedef SEXP (*F2)(SEXP, SEXP);
typedef SEXP (*F3)(SEXP, SEXP, SEXP);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
#define CALL1(n, a) ((F1)fakeR_lookup(n, 1))(a)
#define CALL2(n, a, b) ((F2)fakeR_lookup(n, 2))(a, b)
#define CALL3(n, a, b, c) ((F3)fakeR_lookup(n, 3))(a, b, c)
#define CALL6(n, a, b, c, d, e, f) ((F6)fakeR_lookup(n, 6))(a, b, c, d, e, f)
#define CALL7(n, a, b, c, d, e, f, g) ((F7)fakeR_lookup(n, 7))(a, b, c, d, e, f, g)

static void rd(void* p, size_t sz, size_t n, FILE* f) { if (fread(p, sz, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
static SEXP scalar_int(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
static SEXP matrix(unsigned type, int nr, int nc) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc), d = Rf_allocVector(INTSXP, 2);
    INTEGER(d)[0] = nr; INTEGER(d)[1] = nc;
    Rf_setAttrib(s, R_DimSymbol, d);
    return s;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: run_shim problem.bin result.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hd[8];
    double stdselint;
    rd(hd, 4, 8, f);
    rd(&stdselint, 8, 1, f);
    const int model = hd[0], C = hd[1], m = hd[2], P = hd[3], T = hd[4], N = hd[5], use_map = hd[6], selind = hd[7];
    R_init_genomicMateSelectRb200(NULL);                       /* what R does when the package's shared object is loaded */
    SEXP m_c = Rf_allocVector(INTSXP, C);
    rd(INTEGER(m_c), 4, (size_t)C, f);
    SEXP h = CALL1("gmsr_create", scalar_int(0));
    if (use_map) {
        SEXP cM = Rf_allocVector(REALSXP, m);
        rd(REAL(cM), 8, (size_t)m, f);
        CALL3("gmsr_set_genmap", h, m_c, cM);
    } else {
        SEXP blocks = Rf_allocVector(VECSXP, C);
        for (int c = 0; c < C; ++c) {
            SEXP b = matrix(REALSXP, INTEGER(m_c)[c], INTEGER(m_c)[c]);
            rd(REAL(b), 8, (size_t)INTEGER(m_c)[c] * INTEGER(m_c)[c], f);
            SET_VECTOR_ELT(blocks, c, b);
        }
        CALL2("gmsr_set_recomb_blocks", h, blocks);
    }
    SEXP hap = matrix(INTSXP, 2 * P, m);
    rd(INTEGER(hap), 4, (size_t)2 * P * m, f);
    CALL2("gmsr_set_haplotypes", h, hap);
    SEXP add = matrix(REALSXP, m, T), dom = R_NilValue, asub = R_NilValue;
    rd(REAL(add), 8, (size_t)m * T, f);
    if (model >= 1) { dom = matrix(REALSXP, m, T); rd(REAL(dom), 8, (size_t)m * T, f); }
    if (model == 2) { asub = matrix(REALSXP, m, T); rd(REAL(asub), 8, (size_t)m * T, f); }
    SEXP sire = Rf_allocVector(INTSXP, N), dam = Rf_allocVector(INTSXP, N);
    rd(INTEGER(sire), 4, (size_t)N, f);
    rd(INTEGER(dam), 4, (size_t)N, f);
    SEXP w = R_NilValue;
    if (selind) { w = Rf_allocVector(REALSXP, T); rd(REAL(w), 8, (size_t)T, f); }
    fclose(f);

    SEXP res = CALL7("gmsr_predict", h, scalar_int(model), add, dom, asub, sire, dam);
    SEXP nm = fakeR_names(res);
    if (LENGTH(res) != 2 || LENGTH(nm) != 2 || strcmp(fakeR_string_elt(nm, 0), "out") || strcmp(fakeR_string_elt(nm, 1), "nseg")) {
        fprintf(stderr, "bad result list\n");
        return 2;
    }
    SEXP out = VECTOR_ELT(res, 0), nseg = VECTOR_ELT(res, 1);
    SEXP sel = Rf_allocVector(REALSXP, 1);
    REAL(sel)[0] = stdselint;
    SEXP tidy = CALL6("gmsr_tidy", h, scalar_int(model), scalar_int(T), scalar_int(N), w, sel);
    const int ncomp = model + 1, npairs = T * (T + 1) / 2, nof = model == 0 ? 1 : 2, nslots = T + (selind ? 1 : 0);
    if (XLENGTH(out) != (R_xlen_t)ncomp * npairs * N || XLENGTH(tidy) != (R_xlen_t)4 * nslots * nof * N) { fprintf(stderr, "bad result sizes\n"); return 2; }
    CALL1("gmsr_close", h);
    fakeR_run_finalizer(h);                                    /* the GC would run it again later: must be harmless */

    FILE* g = fopen(argv[2], "wb");
    if (!g) { perror(argv[2]); return 2; }
    int32_t oh[5] = {ncomp, npairs, N, nof, nslots};
    fwrite(oh, 4, 5, g);
    fwrite(REAL(out), 8, (size_t)XLENGTH(out), g);
    fwrite(INTEGER(nseg), 4, (size_t)N, g);
    fwrite(REAL(tidy), 8, (size_t)XLENGTH(tidy), g);
    int32_t depth = fakeR_protect_depth();
    fwrite(&depth, 4, 1, g);
    fclose(g);
    return 0;
}
