/* CPU ORACLE, C restatement (test infrastructure, not product code).
 *
 * Same algorithm as oracle/predcrossvar_oracle.py, i.e. the REFERENCE'S formulation of one cross
 * (R/predCrossVar.R:96-162): drop the non-segregating SNPs (:111-113), materialise the dense
 * m_seg x m_seg cross LD matrix D = calcGameticLD(sire) + calcGameticLD(dam) (:261-282), then one
 * quadratic form per trait pair (R/helpers.R:302), and for modelType "AD" a second one on D*D per
 * trait pair (R/predCrossVar.R:219-222).  It exists so that bench.py can time the reference's CPU
 * algorithm on ALL host cores in bounded time (OpenMP over the rows of D; one cross at a time, like
 * `ncores = 1, nBLASthreads = all` in the reference).  It does not use the delta/het reformulation
 * and it does not skip the zero off-chromosome part of D: like R it touches all m_seg^2 entries.
 * Differences from R, all in R's favour: no gc() calls, no tibble assembly, no name matching, the
 * D*D of :219 is fused into the mat-vec instead of being a fresh allocation.
 *
 * Validated against the NumPy oracle (which is pinned to the vignette outputs) in
 * tests/test_oracle_c.py.  Only tests/, __graft_entry__ and bench.py's cpu legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* recombFreqMat access: one dense column-major m x m matrix, or its diagonal blocks (entries
 * outside the blocks are exactly 0, as in 1-2*genmap2recombfreq()), or the genetic map itself: then an
 * entry is what R/helpers.R:204-213 plus the caller's 1-(2*c1) (vignettes/start_here.Rmd:112) would have
 * stored, evaluated on the fly in the same operation order (for chromosomes whose matrix does not fit). */
typedef struct {
    const double* cM;               /* map positions (with chr), or NULL */
    const int32_t* chr;
    const double* dense;            /* m x m column-major, or NULL */
    int C;                          /* number of blocks */
    const int64_t* off;             /* C+1 block boundaries on the SNP axis */
    const double* const* blocks;    /* blocks[c]: (off[c+1]-off[c])^2 column-major */
    int64_t m;
} rmat;

static inline double r_at(const rmat* R, const int32_t* blk, int64_t j, int64_t k) {
    if (R->cM) {
        if (R->chr[j] != R->chr[k]) return 1.0 - (2.0 * 0.5);            /* c1 = 0.5 across chromosomes  :208-211 */
        const double d = fabs(R->cM[j] - R->cM[k]);                       /* dist(m, "manhattan")          :205 */
        const double c1 = 0.5 * (1.0 - exp(-2.0 * (d / 100.0)));          /*                               :206 */
        return 1.0 - (2.0 * c1);
    }
    if (R->dense) return R->dense[j + R->m * k];
    const int32_t c = blk[j];
    if (c != blk[k]) return 0.0;
    const int64_t o = R->off[c], n = R->off[c + 1] - o;
    return R->blocks[c][(j - o) + n * (k - o)];
}

/* One cross. hap rows are 0/1 bytes of length m. add/dom: m x T column-major.
 * out: npairs x ncomp (pair-major; pairs = T variances then combn order; comps VarA[,VarD]).
 * Returns 0, or -1 on allocation failure. */
int gmsref_pred_one_cross(int64_t m, const uint8_t* s0, const uint8_t* s1, const uint8_t* d0, const uint8_t* d1,
                          const double* Rdense, int C, const int64_t* off, const double* const* blocks,
                          const double* cM, const int32_t* chr,
                          int model_ad, int T, const double* add, const double* dom,
                          double* out, int32_t* nseg_out, int nthreads) {
    rmat R = {cM, chr, Rdense, C, off, blocks, m};
    const int npairs = T * (T + 1) / 2, ncomp = model_ad ? 2 : 1;
    int64_t* seg = (int64_t*)malloc(sizeof(int64_t) * (size_t)m);
    int32_t* blk = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
    if (!seg || !blk) { free(seg); free(blk); return -1; }
    for (int c = 0; c < C && !Rdense && !cM; ++c)
        for (int64_t j = off[c]; j < off[c + 1]; ++j) blk[j] = c;
    int64_t ns = 0;
    for (int64_t j = 0; j < m; ++j) {                       /* R/predCrossVar.R:111-113 */
        const int x = s0[j] + s1[j] + d0[j] + d1[j];
        if (x > 0 && x < 4) seg[ns++] = j;
    }
    *nseg_out = (int32_t)ns;
    memset(out, 0, sizeof(double) * (size_t)npairs * ncomp);
    if (ns == 0) { free(seg); free(blk); return 0; }
    double* D = (double*)malloc(sizeof(double) * (size_t)ns * (size_t)ns);
    double* xs = (double*)malloc(sizeof(double) * (size_t)ns * 6);
    double* v = (double*)malloc(sizeof(double) * (size_t)ns * 2);
    if (!D || !xs || !v) { free(D); free(xs); free(v); free(seg); free(blk); return -1; }
    double *sa = xs, *sb = xs + ns, *da = xs + 2 * ns, *db = xs + 3 * ns, *ps = xs + 4 * ns, *pd = xs + 5 * ns;
    for (int64_t a = 0; a < ns; ++a) {
        const int64_t j = seg[a];
        sa[a] = s0[j]; sb[a] = s1[j]; da[a] = d0[j]; db[a] = d1[j];
        ps[a] = (sa[a] + sb[a]) / 2.0;                      /* colMeans(X)  :263 */
        pd[a] = (da[a] + db[a]) / 2.0;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    /* D = R*((0.5*t(X)%*%X) - p%*%t(p)) for the sire, plus the same for the dam   :264, :281-282 */
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < ns; ++b) {
        double* col = D + (size_t)b * ns;
        for (int64_t a = 0; a < ns; ++a) {
            const double r = r_at(&R, blk, seg[a], seg[b]);
            const double ls = (0.5 * sa[a]) * sa[b] + (0.5 * sb[a]) * sb[b] - ps[a] * ps[b];
            const double ld = (0.5 * da[a]) * da[b] + (0.5 * db[a]) * db[b] - pd[a] * pd[b];
            col[a] = r * ls + r * ld;
        }
    }
    /* per trait pair: quadform(D, a1, a2) and quadform(D*D, d1, d2)   :209-222 ; helpers.R:302 */
    int pr = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int t1 = 0; t1 < T; ++t1)
            for (int t2 = (pass ? t1 + 1 : t1); t2 < (pass ? T : t1 + 1); ++t2, ++pr) {
                const double* a1 = add + (size_t)t1 * m;
                const double* a2 = add + (size_t)t2 * m;
                const double* e1 = model_ad ? dom + (size_t)t1 * m : NULL;
                const double* e2 = model_ad ? dom + (size_t)t2 * m : NULL;
                /* D %*% y  (D is symmetric, so rows of the product are dot products with columns of D) */
#pragma omp parallel for schedule(static)
                for (int64_t a = 0; a < ns; ++a) {
                    const double* col = D + (size_t)a * ns;
                    double sA = 0.0, sD = 0.0;
                    if (model_ad) {
                        for (int64_t b = 0; b < ns; ++b) {
                            const double dv = col[b];
                            sA += dv * a2[seg[b]];
                            sD += (dv * dv) * e2[seg[b]];
                        }
                    } else {
                        for (int64_t b = 0; b < ns; ++b) sA += col[b] * a2[seg[b]];
                    }
                    v[a] = a1[seg[a]] * sA;                  /* x * (D %*% y) */
                    if (model_ad) v[ns + a] = e1[seg[a]] * sD;
                }
                long double tA = 0.0L, tD = 0.0L;            /* colSums accumulates in long double */
                for (int64_t a = 0; a < ns; ++a) { tA += v[a]; if (model_ad) tD += v[ns + a]; }
                out[pr * ncomp] = (double)tA;
                if (model_ad) out[pr * ncomp + 1] = (double)tD;
            }
    free(D); free(xs); free(v); free(seg); free(blk);
    return 0;
}

int gmsref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
