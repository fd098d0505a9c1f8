/* gms_b200.h - C ABI of the B200-native cross-variance engine.
 *
 * This is the drop-in boundary for ONE path of the R package genomicMateSelectR:
 *   predCrossVars -> predOneCross -> calcCrossLD / calcGameticLD -> predOneCrossVar -> quadform
 *   (reference: R/predCrossVar.R:24-282, R/helpers.R:204-213,302), modelType A / AD / DirDom.
 * The reference is pure R and has no FFI for this path; the R glue that a maintainer adds
 * (a .Call shim + a predCrossVars() with the reference's signature) is in r-package/ and is
 * described in INTEGRATION.md.  Everything here is plain C: pointers + sizes, caller-owned
 * host buffers, int status codes, no exceptions across the boundary.
 *
 * Every function returns GMS_OK (0) or a negative GMS_E* code; gms_last_error() gives the text.
 * A handle is bound to ONE CUDA device.  Entry points lock the handle (an internal mutex): calls on one
 * handle from several host threads are serialised, different handles run concurrently (one per GPU).
 * There is NO CPU fallback: without a CUDA device gms_create fails.
 */
#ifndef GMS_B200_H
#define GMS_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define GMS_API __attribute__((visibility("default")))
#else
#define GMS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gms_handle gms_handle;

enum {
    GMS_OK = 0,
    GMS_EINVAL = -1,   /* bad argument (NULL, out-of-range index, non 0/1 haplotype, asymmetric matrix ...) */
    GMS_ECUDA = -2,    /* CUDA runtime error (text in gms_last_error) */
    GMS_ENOMEM = -3,   /* host or device allocation failed */
    GMS_ESTATE = -4    /* call order: map/blocks and haplotypes must be set before predict */
};

/* modelType of predCrossVars / predictCrosses (R/predCrossVar.R:24; R/gmsFunctions.R:683-723). */
enum {
    GMS_MODEL_A = 0,      /* 1 component  per trait pair: VarA                                  */
    GMS_MODEL_AD = 1,     /* 2 components per trait pair: VarA, VarD                            */
    GMS_MODEL_DIRDOM = 2  /* 3 components: VarA, VarD (from add=a, dom=d) and VarBV (from asub) -
                             the two predCrossVars calls of R/gmsFunctions.R:704-723 in one pass */
};

/* Library / build identification, e.g. "gms_b200 0.1 sm_100a". Never NULL. */
GMS_API const char* gms_version(void);

/* Text of the last error on this handle (or of the last failed gms_create if h == NULL). */
GMS_API const char* gms_last_error(const gms_handle* h);

/* Create / destroy an engine on CUDA device `device`. */
GMS_API int gms_create(gms_handle** out, int device);
GMS_API void gms_destroy(gms_handle* h);

/* Optional: run all work of this handle on the caller's CUDA stream (a cudaStream_t passed as
 * void*; NULL = the handle's own stream). Lets a host time the engine with its own events. */
GMS_API int gms_set_stream(gms_handle* h, void* cuda_stream);

/* ---- stage (1): recombination / LD-decay blocks -------------------------------------------
 * Replaces genmap2recombfreq() + the caller's `1-(2*c)` (R/helpers.R:204-213;
 * vignettes/start_here.Rmd:112).  SNPs must be ordered chromosome by chromosome; m_c[c] is the
 * number of SNPs of chromosome c, m = sum(m_c).
 *
 * gms_set_genmap: cM[m] are centimorgan positions; the device builds, per chromosome,
 *   R_c[j,k] = 1-(2*(0.5*(1-exp(-2*(|cM_j-cM_k|/100)))))  and  R_c o R_c  directly in HBM
 *   (the off-chromosome part of the reference's m x m matrix is exactly 0 and is never stored).
 * gms_set_recomb_blocks: the caller already has recombFreqMat (= 1-2c); blocks[c] points to the
 *   c-th diagonal block, column-major m_c[c] x m_c[c] doubles.  Blocks must be symmetric
 *   (the reference documents the matrix as symmetric); asymmetric input -> GMS_EINVAL.
 *   Passing C = 1 with the whole matrix treats the genome as one block (general matrix). */
GMS_API int gms_set_genmap(gms_handle* h, int C, const int64_t* m_c, const double* cM);
GMS_API int gms_set_recomb_blocks(gms_handle* h, int C, const int64_t* m_c, const double* const* blocks);

/* ---- phased haplotypes (haploMat, R/predCrossVar.R:44-46,111-112,262) ----------------------
 * P parents, m SNPs (same order as the map). Parent p owns rows 2p (HapA) and 2p+1 (HapB).
 * _i32cm: R `integer` matrix layout, column-major (2P) x m: element (row r, snp j) at r + 2P*j.
 * _u8rm : row-major uint8 (2P) x m (NumPy layout): element at r*m + j.
 * Values other than 0/1 -> GMS_EINVAL (the reference does not validate). */
GMS_API int gms_set_haplotypes_i32cm(gms_handle* h, int64_t P, int64_t m, const int32_t* hap);
GMS_API int gms_set_haplotypes_u8rm(gms_handle* h, int64_t P, int64_t m, const uint8_t* hap);

/* Optional: evaluation mode.
 * GMS_MODE_DENSE: the dense FP64 tensor (DMMA) contraction against the stored R / RoR blocks; any symmetric recombFreqMat.
 * GMS_MODE_INT8_TENSOR (any symmetric matrix, any model, T <= 32): all contractions run on the int8 tcgen05 tensor cores.
 *   G = (R or RoR) . diag(effects) is rounded, row by row, to 47 bits below a power-of-two ROW scale (>= the row's largest
 *   |entry| / 0.996) and split into 6 balanced radix-256 int8 digit planes; the masks are ternary; int32 accumulation in
 *   TMEM is exact; the digit sums are recombined exactly in 64-bit integers and converted to FP64 once per row. This is
 *   FIXED POINT relative to the row maximum, so its accuracy depends on the data: a SNP with a huge effect that does not
 *   segregate in a cross leaves an absolute error of 2^-48 x (row scale) on terms that may be many orders of magnitude
 *   smaller. The per-cross pass evaluates the crosses in sire order internally (results come back in list order), because
 *   crosses that share a parent share most of their zero mask rows and the kernel skips those.
 *   THE CONTRACT: every table (per parent, per
 *   cross) is followed on the device by a worst-case a-posteriori bound of that error (csrc/verify.cu); a unit passes when
 *   the bound is at most tau x sqrt(X[t,t] X[s,s]) for every trait pair (tau = 2^-31 ~ 4.7e-10 by default,
 *   gms_set_int8_tolerance), i.e. every variance is good to tau relative and every covariance to tau of its variance scale. Units that fail (and units with
 *   >= 16384 non-zero mask entries on one chromosome, where the 64-bit digit recombination could overflow) are
 *   re-evaluated by the FP64 contraction inside the same call, at most 16384 per table in-stream; if more fail, the whole
 *   call is redone with GMS_MODE_DENSE (gms_stats.last_fallback_units / last_fallback_overflow report both). Results
 *   therefore meet the 1e-9 contract for ANY finite input; only the speed is data dependent.
 * GMS_MODE_AUTO (default): GMS_MODE_INT8_TENSOR when its operand images fit in HBM (a few GB at cassava scale), else
 *   GMS_MODE_DENSE.
 * GMS_MODE_MAP_SCAN: the map-structured fast path: when the blocks came from gms_set_genmap() with positions
 *   sorted within each chromosome, R[j,k] = exp(-2|x_j-x_k|/100) factorises and every quadratic form becomes a prefix-sum
 *   scan over the non-zero SNPs of a unit (O(m T^2) per cross instead of O(m^2 T)); same results to ~1e-13. Returns
 *   GMS_ESTATE if the handle's blocks did not come from a sorted map. */
enum { GMS_MODE_DENSE = 0, GMS_MODE_MAP_SCAN = 1, GMS_MODE_INT8_TENSOR = 2, GMS_MODE_AUTO = 3 };
GMS_API int gms_set_mode(gms_handle* h, int mode);
/* tau of the int8 error bound (see above). tau <= 0 switches the verification off (benchmarks of its cost only). */
GMS_API int gms_set_int8_tolerance(gms_handle* h, double tau);

/* Per-handle cache (default on). The handle remembers a hash of the last effects: a later call with the same effects,
 * model, T and mode reuses the operand images in HBM and every per-parent table row it already has, and computes rows
 * only for parents it has not seen (mate-selection loops evaluate many candidate cross lists against one set of marker
 * effects; R/predCrossVar.R:44-46 subsets haploMat per call instead). enabled = 0: every call recomputes everything. */
GMS_API int gms_set_cache(gms_handle* h, int enabled);

/* Optional chromosome shard: this handle only sums over chromosomes [c_begin, c_end).  Outputs of
 * gms_predict are then PARTIAL sums (variances and Nsegsnps are additive over chromosomes); the
 * host adds the partial slabs of the shards (NCCL all-reduce in the multi-GPU driver).
 * Default: all chromosomes. */
GMS_API int gms_set_chromosome_shard(gms_handle* h, int c_begin, int c_end);

/* Number of trait pairs T(T+1)/2 and their order: the T variances (t,t) first, then
 * combn(traits,2) order (0,1),(0,2),...,(T-2,T-1)  (R/predCrossVar.R:128-138).
 * pair_t1/pair_t2 (length npairs, may be NULL) receive the 0-based trait indices. */
GMS_API int gms_trait_pairs(int T, int32_t* pair_t1, int32_t* pair_t2);

/* ---- the hot path -------------------------------------------------------------------------
 * gms_predict = predCrossVars() for N crosses in one call.
 *   model : GMS_MODEL_*
 *   T     : number of traits (1..32)
 *   add   : m x T column-major (trait t at add + t*m): AddEffectList     (R/predCrossVar.R:32-43)
 *   dom   : m x T, DomEffectList, required for AD and DIRDOM, else NULL
 *   asub  : m x T, allele-substitution effects for DIRDOM's VarBV, else NULL
 *           (all effects must be finite: NA / NaN / Inf -> GMS_EINVAL; the reference would propagate them)
 *   N, sire[N], dam[N] : 0-based parent indices of each cross (selfs and duplicates allowed)
 *   out   : N x npairs x ncomp doubles, cross-major then pair-major; component order
 *           VarA [, VarD [, VarBV]]; pair order of gms_trait_pairs
 *   nseg  : N int32, Nsegsnps = #{SNP: 0 < sum of the 4 parental haplotypes < 4}
 *           (R/predCrossVar.R:111-113).  Crosses with nseg == 0 get zeros; the R glue drops
 *           them like the reference's list()/unnest path does (R/predCrossVar.R:156-160).
 * Host buffers only; the call copies inputs to HBM, runs, copies results back, and returns. */
GMS_API int gms_predict(gms_handle* h, int model, int T,
                const double* add, const double* dom, const double* asub,
                int64_t N, const int32_t* sire, const int32_t* dam,
                double* out, int32_t* nseg);

/* Multi-GPU from ONE host process (what a single R session needs): `hs[0..nh)` are handles on different
 * devices that were given the same map / blocks and haplotypes. The cross list is cut into nh contiguous
 * shards, one host thread per handle runs gms_predict on its shard and writes straight into the caller's
 * `out` / `nseg` slices - no collective, only this implicit gather. Returns the first non-zero status
 * (its text is copied to hs[0]'s gms_last_error). */
GMS_API int gms_predict_sharded(gms_handle* const* hs, int nh, int model, int T,
                                const double* add, const double* dom, const double* asub,
                                int64_t N, const int32_t* sire, const int32_t* dam,
                                double* out, int32_t* nseg);

/* The same work split in three so that a benchmark can time the device part alone with inputs
 * resident in HBM: gms_stage_inputs (host->device), gms_compute (kernels only, asynchronous on
 * the handle's stream unless `sync` != 0), gms_fetch_outputs (device->host, synchronises). */
GMS_API int gms_stage_inputs(gms_handle* h, int model, int T,
                     const double* add, const double* dom, const double* asub,
                     int64_t N, const int32_t* sire, const int32_t* dam);
GMS_API int gms_compute(gms_handle* h, int sync);
GMS_API int gms_fetch_outputs(gms_handle* h, double* out, int32_t* nseg);

/* ---- multi-GPU, one process per GPU -----------------------------------------------
 * Every rank passes the WHOLE cross list and its own contiguous range [x_begin, x_end) of it; outputs have
 * x_end - x_begin rows. With gms_set_parent_shard(h, rank, nranks) the per-parent stage is split as well: the parents
 * of the whole list (the same on every rank) are cut into nranks contiguous slices, gms_compute_parents evaluates this
 * rank's slice and packs it into an exchange buffer, the host all-gathers the slices (one ncclAllGather of
 * doubles_per_part doubles per rank, in place in *dev_buf; NULL / 0 when nothing had to be computed), and
 * gms_compute_crosses unpacks them and runs the per-cross stage. gms_compute = both phases on an unsharded handle. */
GMS_API int gms_stage_inputs_range(gms_handle* h, int model, int T,
                     const double* add, const double* dom, const double* asub,
                     int64_t N, const int32_t* sire, const int32_t* dam, int64_t x_begin, int64_t x_end);
GMS_API int gms_set_parent_shard(gms_handle* h, int index, int parts);
GMS_API int gms_compute_parents(gms_handle* h);
GMS_API int gms_parent_exchange(gms_handle* h, void** dev_buf, int64_t* doubles_per_part, int32_t* parts);
GMS_API int gms_compute_crosses(gms_handle* h, int sync);
/* Device addresses of the result slab (N x npairs x ncomp doubles) and of Nsegsnps (N int32) of the last compute, valid
 * until the next stage call: lets a multi-GPU host reduce chromosome-sharded partial slabs in place (ncclAllReduce)
 * before gms_fetch_outputs copies them out. Work queued on the handle's stream may still be writing them. */
GMS_API int gms_device_outputs(gms_handle* h, void** out_dev, void** nseg_dev, int64_t* n_rows);

/* Cross-validation batching (R/parentWiseCrossVal.R:430-490 calls the path once per (Repeat, Fold) on the same haploMat,
 * recombFreqMat and - per fold - the same crosses): nsets effect sets against ONE cross list in one call.
 * add / dom / asub: nsets consecutive m x T column-major matrices; out: nsets consecutive N x npairs x ncomp slabs;
 * nseg: N (the same for every set). Cross-side device state (cross lists, mask images) is staged once. */
GMS_API int gms_predict_batched(gms_handle* h, int model, int nsets, int T,
                     const double* add, const double* dom, const double* asub,
                     int64_t N, const int32_t* sire, const int32_t* dam, double* out, int32_t* nseg);

/* ---- "next" rows of the scope table ---------------------------------------------
 * gms_selind: selection-index variance w'Gw per cross and component from an `out` slab
 *   (R/gmsFunctions.R:825-853).  w[T]; si[N x ncomp].  Host-side O(N T^2).
 * gms_cross_means: predCrossMeans (R/predCrossVar.R:336-399) on the device-resident haplotypes
 *   (dosage = HapA + HapB): parent GEBV = dosage . add_t, mean BV = (sire+dam)/2;
 *   if dom != NULL also the Falconer-MacKay eqn 14.6 TGV mean.  mean_bv, mean_tgv: N x T
 *   cross-major (mean_tgv may be NULL iff dom is NULL); gebv: P x T parent-major or NULL. */
GMS_API int gms_selind(int model, int T, int64_t N, const double* out, const double* w, double* si);
/* gms_tidy: the tidy table of predictCrosses() (R/gmsFunctions.R:726-871) for the crosses of the last gms_compute /
 *   gms_predict on this handle, computed on the device from the result slab and the effects already in HBM:
 *   tidy[x][predOf][slot][4] = (predMean, predVar, predSD, predUsefulness = predMean + std_sel_int * predSD);
 *   predOf: BV [, TGV]  (1 for model A, else 2; TGV = VarBV + VarDD for AD, VarA + VarD for DirDom, :772-800);
 *   slot: [SELIND (only if si_weights != NULL; w'Gw and w'mean, :805-853),] trait 0 .. T-1.
 *   Means: MeanBV = mid-parent GEBV from the allele-substitution effects (add for A / AD, asub for DirDom);
 *   MeanTGV = MeanBV for AD (:737-741), Falconer-MacKay 14.6 from (add, dom) for DirDom (R/predCrossVar.R:379-396).
 *   The model fixes which effects predictCrosses passes, so no effect arrays are needed here. */
GMS_API int gms_tidy(gms_handle* h, const double* si_weights, double std_sel_int, double* tidy);
GMS_API int gms_cross_means(gms_handle* h, int T, const double* add, const double* dom,
                    int64_t N, const int32_t* sire, const int32_t* dam,
                    double* gebv, double* mean_bv, double* mean_tgv);

/* ---- introspection (benchmarks, tests) ---------------------------------------------------- */
typedef struct gms_stats {
    int64_t m;              /* SNPs                                                         */
    int64_t m_padded;       /* padded SNP axis (each chromosome rounded up to 128)          */
    int64_t S2;             /* sum_c m_c^2 over this handle's chromosome shard              */
    int64_t tile_bytes;     /* HBM bytes of the R and RoR tile images                       */
    int32_t C;              /* chromosomes                                                  */
    int32_t P;              /* parents                                                      */
    int32_t last_T;
    int32_t last_model;
    int64_t last_N;
    int64_t last_parents_used;    /* distinct parents in the last cross list               */
    int32_t last_launches;        /* kernels launched by the last gms_compute              */
    int32_t last_mode;            /* GMS_MODE_* the last gms_compute actually ran (AUTO resolved) */
    float   last_ms_parent_stage; /* per-parent contractions (device time, CUDA events)    */
    float   last_ms_cross_stage;  /* per-cross dominance contraction + its finalisation     */
    float   last_ms_cross_kernel; /* the per-cross contraction kernel alone (dominant kernel) */
    float   last_ms_assemble;     /* finalize + assemble + Nsegsnps                        */
    float   last_ms_total;        /* whole gms_compute                                     */
    double  last_flops_cross;     /* FP64 flops EXECUTED by the cross contraction launch   */
    double  last_flops_parent;    /* FP64 flops executed by the per-parent launches        */
    int64_t last_fallback_units;  /* units (parents + crosses) that failed the int8 error bound and were re-evaluated in FP64 */
    int64_t last_parents_computed;/* per-parent table rows this handle computed in the last call (cache, parent shard) */
    int32_t last_fallback_overflow; /* 1: more units failed than the in-stream lists hold; the call was redone in GMS_MODE_DENSE */
    int32_t last_cache_hit;       /* 1: same effects as the previous call: operand images and cached table rows reused */
} gms_stats;
GMS_API int gms_get_stats(gms_handle* h, gms_stats* s);

#ifdef __cplusplus
}
#endif
#endif /* GMS_B200_H */
