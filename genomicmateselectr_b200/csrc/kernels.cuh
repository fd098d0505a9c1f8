// Launchers of the small (HBM-bound) kernels around the contraction; see kernels.cu.
#pragma once
#include "common.cuh"
#include "i8digits.cuh"

namespace gms {

// stage (1): LD-decay tile images (R and RoR) in HBM, straight from the genetic map or from a
// caller-made recombFreqMat block. `asym_count` (device int) counts asymmetric entries.
cudaError_t launch_build_tiles_from_map(const double* cM_dev, const ChromDesc& cd, double* R, double* R2,
                                        cudaStream_t st);
cudaError_t launch_build_tiles_from_block(const double* blk_dev, const ChromDesc& cd, double* R, double* R2,
                                          int* asym_count, cudaStream_t st);

// haplotype bit-planes on the padded SNP axis. `bad_count` counts values outside {0,1}.
cudaError_t launch_pack_i32cm(const int32_t* hap_dev, long long P, long long m, const ChromDesc* chrom, int C,
                              unsigned* planeA, unsigned* planeB, long long wpp, int* bad_count, cudaStream_t st);
cudaError_t launch_pack_u8rm(const uint8_t* hap_dev, long long P, long long m, const ChromDesc* chrom, int C,
                             unsigned* planeA, unsigned* planeB, long long wpp, int* bad_count, cudaStream_t st);

// effects: caller layout (m x T column-major, unpadded) -> padded emat (T x mp_total)
cudaError_t launch_scatter_effects(const double* eff_dev, long long m, int T, const ChromDesc* chrom, int C,
                                   double* emat, long long mp_total, cudaStream_t st);

// X[n][t][s] = sum_split Z[split][n][t][s] + Z[split][n][s][t]   (sym)   or   sum_split Z[split][n][t][s]
// row_index: unit u goes to row row_index[u] of X; n_dev: device-side unit count (<= n); zcap: unit capacity of Z (default n)
cudaError_t launch_finalize_sym(const double* Z, int nsplit, long long n, int T, double* X, cudaStream_t st,
                                bool sym = true, const int* row_index = nullptr, const int* n_dev = nullptr,
                                long long zcap = 0);

// per-parent table rows <-> exchange buffer [row][class][T*T] (rows listed by parent id in `ids`)
cudaError_t launch_pack_rows(const double* const* tabs, int ncls, int TT, const int* ids, long long n, double* buf, cudaStream_t st);
cudaError_t launch_unpack_rows(const double* buf, int ncls, int TT, const int* ids, long long n, double* const* tabs, cudaStream_t st);

// ---- map-structured fast path (scan.cu) ----
struct ScanParams {
    const ChromDesc* chrom;
    const int* split_begin;      // [nsplit + 1] ranges of chromosomes
    const double* S;             // [mp][3][T]: e, a e, b e
    const unsigned* planeA;
    const unsigned* planeB;
    long long wpp;
    const int* unit_a;
    const int* unit_b;
    long long n_units;
    int T;
    int mode;
    double* Xpart;               // [nsplit][n_units][T*T], symmetric
};
cudaError_t launch_scan_prepare(const double* emat, const double* apos, const double* bpos, long long mp_total, int T,
                                double* S, cudaStream_t st);
cudaError_t launch_scan(const ScanParams& p, int nsplit, cudaStream_t st);

// ---- int8 tensor-core path for the quadratic-form tables (i8path.cu) ----
struct I8Params {
    const ChromDesc* chrom;
    const RowTile* jtiles;       // (c, J): 64-row tiles of the lower block triangle, ordered by (c, J)
    const int* split_begin;      // [nsplit + 1] ranges of jtiles
    const signed char* xi;       // [X tile][global 64-k block][8 KiB image]
    long long nkb_total;         // mp_total / 64
    const signed char* G;        // digit images
    const long long* gbase;      // [C] byte offset of each chromosome's digit images
    const double* scale;         // [T][mp_total] power-of-two row scales
    const double* emat;          // [T][mp_total] effects that multiply the rows (d_t[j])
    long long mp_total;
    const unsigned* planeA;
    const unsigned* planeB;
    long long wpp;
    const int* unit_a;
    const int* unit_b;
    long long n_units;
    int T;
    int mode;
    double* Zpart;               // [nsplit][n_units][T*T]
    int stages;                  // pipeline depth (set by launch_i8_contract)
    int debug;                   // -DGMS_PROFILE builds only (GMS_I8_DEBUG bit mask; results are garbage): 1 no epilogue arithmetic, 2 no MMAs
};
size_t i8_smem_bytes(int T, bool cg2);
cudaError_t launch_i8_build_xi(const unsigned* planeA, const unsigned* planeB, long long wpp, const int* sire,
                               const int* dam, long long n_units, int mode, long long nkb_total, signed char* xi,
                               cudaStream_t st);
cudaError_t launch_i8_build_digits(const double* tiles, const ChromDesc* chrom_d, int c_begin, int c_end, int C,
                                   const int* blkpref_d, int nblks, const double* emat, long long mp_total, int T,
                                   double* scale, double* smax, const long long* gbase_d, signed char* G, int layout,
                                   cudaStream_t st);     // layout: 0 / 1 = plain / cta_group::2; smax[T][C]: see verify.cu
cudaError_t launch_i8_contract(const I8Params& p, int ntiles, int nsplit, bool pair, cudaStream_t st);
long long i8_block_bytes();         // bytes of the digit image of one (64 x 64 block, trait)
int i8_npart(int T);        // partial slabs per split written by launch_i8_contract
int i8_max_tiles_per_split();       // the plan must cut the 64-row tile list into splits no longer than this

// ---- accuracy contract of the int8 path (verify.cu): a-posteriori error bound per unit, list of failing units ----
struct VerifyParams {
    const ChromDesc* chrom;
    int c_begin, c_end, C;
    const unsigned* planeA;
    const unsigned* planeB;
    long long wpp;
    const int* unit_a;
    const int* unit_b;
    long long n_units;
    int mode;                    // MASK_*
    const double* emat;          // [T][mp_total] effects of the class
    long long mp_total;
    int T;
    const double* smax;          // [T][C] largest row scale per trait and chromosome (0: all rows zero)
    const double* table;         // [rows][T][T] symmetric table the units were written to
    int rows_by_parent;          // table row of unit u: unit_a[u] (per-parent tables) or u (per-cross table)
    double tau;                  // tolerance on |error| / sqrt(X[t,t] X[s,s])
    int* counter;                // number of failing units (may exceed cap)
    int cap;                     // capacity of the lists
    int* list_a;                 // failing units: parents, table row
    int* list_b;
    int* list_row;
    double* bound_out;           // optional [n_units]: the bound itself (tests, statistics)
    double* part_b;              // scratch [chromosomes of the shard][n_units][T]: per-chromosome partial sums of B_t
    int* part_n;                 // scratch [chromosomes of the shard][n_units]: non-zero mask entries per chromosome
};
size_t i8_verify_scratch_doubles(long long n_units, int nchrom, int T);
cudaError_t launch_i8_smax(const double* rowmax, const ChromDesc* chrom_d, int c_begin, int c_end, int C, long long mp_total,
                           int T, double* smax, cudaStream_t st);
cudaError_t launch_i8_verify(const VerifyParams& p, cudaStream_t st);

// Nsegsnps per cross over words [w_begin, w_end) of the padded axis.
cudaError_t launch_nseg(const unsigned* planeA, const unsigned* planeB, long long wpp, long long w_begin,
                        long long w_end, const int* sire, const int* dam, long long N, int* nseg, cudaStream_t st);

// out[x][pair][comp] from the per-parent tables and the per-cross table.
struct AssembleParams {
    const double* QA;    // [P][T][T] additive table of (add, R, delta), indexed by parent id
    const double* PD;    // [P][T][T] dominance per-parent table of (dom, RoR, het) or NULL
    const double* QB;    // [P][T][T] additive table of asub (DirDom VarBV) or NULL
    const double* X;     // [N][T][T] per-cross dominance term or NULL
    const int* sire;     // [N] parent ids
    const int* dam;
    const int* pair_t1;  // [npairs]
    const int* pair_t2;
    long long N;
    int T, npairs, ncomp;
    double* out;         // [N][npairs][ncomp]
};
cudaError_t launch_assemble(const AssembleParams& p, cudaStream_t st);

// predictCrosses() tidy table on the device: tidy[x][predOf][trait slot][predMean, predVar, predSD, predUsefulness]
struct TidyParams {
    const double* out;       // [N][npairs][ncomp] result slab
    const double* gebv;      // [P][T] parent GEBVs from the allele-substitution effects
    const double* mean_tgv;  // [N][T] Falconer-MacKay TGV means (DirDom) or NULL
    const int* sire;
    const int* dam;
    const int* pair_t1;
    const int* pair_t2;
    const double* w;         // [T] selection-index weights (device) or NULL
    double std_sel_int;
    long long N;
    int T, npairs, ncomp, model;
    double* tidy;
};
cudaError_t launch_tidy(const TidyParams& p, cudaStream_t st);

// predCrossMeans: gebv[p][t] = sum_j dose_pj add_t[j]; TGV mean per cross (Falconer-MacKay 14.6).
cudaError_t launch_gebv(const unsigned* planeA, const unsigned* planeB, long long wpp, const int* plist,
                        long long nP, const double* emat, long long mp_total, int T, double* gebv, cudaStream_t st);
cudaError_t launch_tgv_mean(const unsigned* planeA, const unsigned* planeB, long long wpp, const int* sire,
                            const int* dam, long long N, const double* emat_add, const double* emat_dom,
                            long long mp_total, int T, double* mean_tgv, cudaStream_t st);

}  // namespace gms
