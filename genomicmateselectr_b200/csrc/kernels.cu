// HBM-bound kernels around the contraction: LD-decay tile builders (stage 1 of the path),
// haplotype bit-packing, effect scatter, table finalisation, per-cross assembly, Nsegsnps, means.
#include "kernels.cuh"

namespace gms {

// ------------------------------------------------------------------------------------------
// Tile images. One block per 32 KiB chunk; element order documented in common.cuh.
__device__ __forceinline__ void chunk_to_tile(long long ch, int& I, int& kc) {
    int i = (int)((sqrt(1.0 + 2.0 * (double)ch) - 1.0) * 0.5);
    while (chunks_before_rowtile(i + 1) <= ch) ++i;
    while (chunks_before_rowtile(i) > ch) --i;
    I = i;
    kc = (int)(ch - chunks_before_rowtile(i));
}

__device__ __forceinline__ void elem_to_rowcol(int e, int I, int kc, int& row, int& col) {
    const int hi = e & 1, lane = (e >> 1) & 31, mb = (e >> 6) & 7, q = e >> 9;
    row = I * BM + mb * 16 + (lane >> 2) + 8 * hi;
    col = kc * BK + 4 * q + (lane & 3);
}

// R/helpers.R:204-213 + `1-(2*c1)` (vignettes/start_here.Rmd:112), same operation order as R.
__global__ void build_tiles_from_map_kernel(const double* __restrict__ cM, int m, long long tile_off,
                                            double* __restrict__ R, double* __restrict__ R2) {
    int I, kc;
    chunk_to_tile(blockIdx.x, I, kc);
    double* r_out = R + tile_off + (long long)blockIdx.x * CHUNK;
    double* r2_out = R2 + tile_off + (long long)blockIdx.x * CHUNK;
    const bool diag = (kc >> 2) == I;
    for (int e = threadIdx.x; e < CHUNK; e += blockDim.x) {
        int row, col;
        elem_to_rowcol(e, I, kc, row, col);
        double r = 0.0;
        if (row < m && col < m) {
            const double d = fabs(cM[row] - cM[col]);
            const double c1 = 0.5 * (1.0 - exp(-2.0 * (d / 100.0)));
            r = 1.0 - (2.0 * c1);
        }
        double r2 = r * r;                       // D*D of R/predCrossVar.R:219 up to exact powers of two
        if (diag) { r *= 0.5; r2 *= 0.5; }
        r_out[e] = r;
        r2_out[e] = r2;
    }
}

__global__ void build_tiles_from_block_kernel(const double* __restrict__ blk, int m, long long tile_off,
                                              double* __restrict__ R, double* __restrict__ R2,
                                              int* __restrict__ asym_count) {
    int I, kc;
    chunk_to_tile(blockIdx.x, I, kc);
    double* r_out = R + tile_off + (long long)blockIdx.x * CHUNK;
    double* r2_out = R2 + tile_off + (long long)blockIdx.x * CHUNK;
    const bool diag = (kc >> 2) == I;
    int bad = 0;
    for (int e = threadIdx.x; e < CHUNK; e += blockDim.x) {
        int row, col;
        elem_to_rowcol(e, I, kc, row, col);
        double r = 0.0;
        if (row < m && col < m) {
            r = blk[(long long)row + (long long)m * col];
            const double rt = blk[(long long)col + (long long)m * row];
            const double scale = fmax(1.0, fmax(fabs(r), fabs(rt)));
            if (!(fabs(r - rt) <= 1e-12 * scale)) ++bad;    // also catches NaN
        }
        double r2 = r * r;
        if (diag) { r *= 0.5; r2 *= 0.5; }
        r_out[e] = r;
        r2_out[e] = r2;
    }
    if (bad) atomicAdd(asym_count, bad);
}

cudaError_t launch_build_tiles_from_map(const double* cM_dev, const ChromDesc& cd, double* R, double* R2,
                                        cudaStream_t st) {
    const long long nch = chunks_of_chrom(cd.mp / BM);
    build_tiles_from_map_kernel<<<(unsigned)nch, 256, 0, st>>>(cM_dev, cd.m, cd.tile_off, R, R2);
    return cudaGetLastError();
}
cudaError_t launch_build_tiles_from_block(const double* blk_dev, const ChromDesc& cd, double* R, double* R2,
                                          int* asym_count, cudaStream_t st) {
    const long long nch = chunks_of_chrom(cd.mp / BM);
    build_tiles_from_block_kernel<<<(unsigned)nch, 256, 0, st>>>(blk_dev, cd.m, cd.tile_off, R, R2, asym_count);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Haplotype bit-planes. Word w covers padded indices [32w, 32w+32); chromosomes start on
// multiples of 128 so a word never straddles two chromosomes.
__device__ __forceinline__ bool word_to_snp(long long w, const ChromDesc* chrom, int C, long long& snp0, int& nvalid) {
    const long long pidx = w * 32;
    for (int c = 0; c < C; ++c) {
        const ChromDesc cd = chrom[c];
        if (pidx >= cd.pad_off && pidx < (long long)cd.pad_off + cd.mp) {
            const long long local = pidx - cd.pad_off;
            snp0 = cd.snp_off + local;
            const long long left = (long long)cd.m - local;
            nvalid = left <= 0 ? 0 : (left >= 32 ? 32 : (int)left);
            return true;
        }
    }
    nvalid = 0;
    snp0 = 0;
    return false;
}

__global__ void pack_i32cm_kernel(const int32_t* __restrict__ hap, long long P, long long m,
                                  const ChromDesc* __restrict__ chrom, int C, unsigned* __restrict__ planeA,
                                  unsigned* __restrict__ planeB, long long wpp, int* __restrict__ bad_count) {
    const long long npb = (P + blockDim.x - 1) / blockDim.x;          // 1-D grid over (word, parent block)
    const long long w = blockIdx.x / npb;
    const long long p = (blockIdx.x - w * npb) * blockDim.x + threadIdx.x;
    if (p >= P) return;
    long long snp0; int nvalid;
    word_to_snp(w, chrom, C, snp0, nvalid);
    unsigned a = 0, b = 0; int bad = 0;
    for (int i = 0; i < nvalid; ++i) {
        const long long j = snp0 + i;
        const int va = hap[2 * p + 2 * P * j], vb = hap[2 * p + 1 + 2 * P * j];
        bad += ((unsigned)va > 1u) + ((unsigned)vb > 1u);
        a |= (unsigned)(va & 1) << i;
        b |= (unsigned)(vb & 1) << i;
    }
    planeA[p * wpp + w] = a;
    planeB[p * wpp + w] = b;
    if (bad) atomicAdd(bad_count, bad);
}

__global__ void pack_u8rm_kernel(const uint8_t* __restrict__ hap, long long P, long long m,
                                 const ChromDesc* __restrict__ chrom, int C, unsigned* __restrict__ planeA,
                                 unsigned* __restrict__ planeB, long long wpp, int* __restrict__ bad_count) {
    const long long nwb = (wpp + blockDim.x - 1) / blockDim.x;       // 1-D grid over (parent, word block): P is not limited by gridDim.y
    const long long p = blockIdx.x / nwb;
    const long long w = (blockIdx.x - p * nwb) * blockDim.x + threadIdx.x;
    if (w >= wpp) return;
    long long snp0; int nvalid;
    word_to_snp(w, chrom, C, snp0, nvalid);
    const uint8_t* ra = hap + (2 * p) * m + snp0;
    const uint8_t* rb = hap + (2 * p + 1) * m + snp0;
    unsigned a = 0, b = 0; int bad = 0;
    for (int i = 0; i < nvalid; ++i) {
        const unsigned va = ra[i], vb = rb[i];
        bad += (va > 1u) + (vb > 1u);
        a |= (va & 1u) << i;
        b |= (vb & 1u) << i;
    }
    planeA[p * wpp + w] = a;
    planeB[p * wpp + w] = b;
    if (bad) atomicAdd(bad_count, bad);
}

cudaError_t launch_pack_i32cm(const int32_t* hap_dev, long long P, long long m, const ChromDesc* chrom, int C,
                              unsigned* planeA, unsigned* planeB, long long wpp, int* bad_count, cudaStream_t st) {
    const long long blocks = ((P + 127) / 128) * wpp;
    if (blocks > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    pack_i32cm_kernel<<<(unsigned)blocks, 128, 0, st>>>(hap_dev, P, m, chrom, C, planeA, planeB, wpp, bad_count);
    return cudaGetLastError();
}
cudaError_t launch_pack_u8rm(const uint8_t* hap_dev, long long P, long long m, const ChromDesc* chrom, int C,
                             unsigned* planeA, unsigned* planeB, long long wpp, int* bad_count, cudaStream_t st) {
    const long long blocks = ((wpp + 127) / 128) * P;
    if (blocks > 0x7FFFFFFFll) return cudaErrorInvalidConfiguration;
    pack_u8rm_kernel<<<(unsigned)blocks, 128, 0, st>>>(hap_dev, P, m, chrom, C, planeA, planeB, wpp, bad_count);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
__global__ void scatter_effects_kernel(const double* __restrict__ eff, long long m, int T,
                                       const ChromDesc* __restrict__ chrom, int C, double* __restrict__ emat,
                                       long long mp_total) {
    const long long pidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (pidx >= mp_total) return;
    double v = 0.0;
    for (int c = 0; c < C; ++c) {
        const ChromDesc cd = chrom[c];
        if (pidx >= cd.pad_off && pidx < (long long)cd.pad_off + cd.mp) {
            const long long local = pidx - cd.pad_off;
            if (local < cd.m) v = eff[(long long)t * m + cd.snp_off + local];
            break;
        }
    }
    emat[(long long)t * mp_total + pidx] = v;
}
cudaError_t launch_scatter_effects(const double* eff_dev, long long m, int T, const ChromDesc* chrom, int C,
                                   double* emat, long long mp_total, cudaStream_t st) {
    dim3 grid((unsigned)((mp_total + 255) / 256), (unsigned)T);
    scatter_effects_kernel<<<grid, 256, 0, st>>>(eff_dev, m, T, chrom, C, emat, mp_total);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Z partial slabs are [nsplit][zcap][T*T] (zcap units of capacity, n <= zcap of them live: n itself or *n_dev).
// row_index: unit u is written to table row row_index[u] (per-parent tables are indexed by parent id; the FP64
// re-evaluation of units that failed the int8 bound writes back to the rows verify.cu listed).
__global__ void finalize_sym_kernel(const double* __restrict__ Z, int nsplit, long long zcap, long long n,
                                    const int* __restrict__ n_dev, const int* __restrict__ row_index, int T,
                                    double* __restrict__ X, bool sym) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int TT = T * T;
    if (n_dev) n = min((long long)*n_dev, n);
    if (idx >= n * TT) return;
    const long long u = idx / TT;
    const int rem = (int)(idx - u * TT);
    const int t = rem / T, s = rem - t * T;
    double v = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) {           // fixed order: bit-reproducible
        const double* z = Z + ((long long)sp * zcap + u) * TT;
        v += sym ? z[t * T + s] + z[s * T + t] : z[t * T + s];
    }
    const long long row = row_index ? (long long)row_index[u] : u;
    X[row * TT + rem] = v;
}
cudaError_t launch_finalize_sym(const double* Z, int nsplit, long long n, int T, double* X, cudaStream_t st, bool sym,
                                const int* row_index, const int* n_dev, long long zcap) {
    const long long total = n * T * T;
    if (total == 0) return cudaSuccess;
    finalize_sym_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Z, nsplit, zcap > 0 ? zcap : n, n, n_dev, row_index, T, X, sym);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Exchange of per-parent table rows between ranks (parent-sharded per-parent stage): rows of up to 3 tables <-> one
// contiguous buffer [row i][class k][T*T], i = position in the list of parents being computed.
struct RowTabs { double* t[3]; };
__global__ void pack_rows_kernel(RowTabs tabs, int ncls, int TT, const int* __restrict__ ids, long long n, double* __restrict__ buf) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * ncls * TT) return;
    const long long i = idx / (ncls * TT);
    const int rem = (int)(idx - i * (ncls * TT)), k = rem / TT, e = rem - k * TT;
    buf[idx] = tabs.t[k][(long long)ids[i] * TT + e];
}
__global__ void unpack_rows_kernel(const double* __restrict__ buf, int ncls, int TT, const int* __restrict__ ids, long long n, RowTabs tabs) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * ncls * TT) return;
    const long long i = idx / (ncls * TT);
    const int rem = (int)(idx - i * (ncls * TT)), k = rem / TT, e = rem - k * TT;
    tabs.t[k][(long long)ids[i] * TT + e] = buf[idx];
}
cudaError_t launch_pack_rows(const double* const* tabs, int ncls, int TT, const int* ids, long long n, double* buf, cudaStream_t st) {
    const long long total = n * ncls * TT;
    if (total == 0) return cudaSuccess;
    RowTabs rt{{const_cast<double*>(tabs[0]), const_cast<double*>(tabs[1]), const_cast<double*>(tabs[2])}};
    pack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rt, ncls, TT, ids, n, buf);
    return cudaGetLastError();
}
cudaError_t launch_unpack_rows(const double* buf, int ncls, int TT, const int* ids, long long n, double* const* tabs, cudaStream_t st) {
    const long long total = n * ncls * TT;
    if (total == 0) return cudaSuccess;
    RowTabs rt{{tabs[0], tabs[1], tabs[2]}};
    unpack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(buf, ncls, TT, ids, n, rt);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Nsegsnps (R/predCrossVar.R:111-113): x = sum of the 4 haplotype rows, keep 0 < x < 4.
__global__ void nseg_kernel(const unsigned* __restrict__ planeA, const unsigned* __restrict__ planeB, long long wpp,
                            long long w_begin, long long w_end, const int* __restrict__ sire,
                            const int* __restrict__ dam, long long N, int* __restrict__ nseg) {
    const long long x = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (x >= N) return;
    const unsigned* As = planeA + (long long)sire[x] * wpp;
    const unsigned* Bs = planeB + (long long)sire[x] * wpp;
    const unsigned* Ad = planeA + (long long)dam[x] * wpp;
    const unsigned* Bd = planeB + (long long)dam[x] * wpp;
    int cnt = 0;
    for (long long w = w_begin + lane; w < w_end; w += 32) {
        const unsigned a = As[w], b = Bs[w], c = Ad[w], d = Bd[w];
        cnt += __popc((a | b | c | d) & ~(a & b & c & d));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) nseg[x] = cnt;
}
cudaError_t launch_nseg(const unsigned* planeA, const unsigned* planeB, long long wpp, long long w_begin,
                        long long w_end, const int* sire, const int* dam, long long N, int* nseg, cudaStream_t st) {
    if (N == 0) return cudaSuccess;
    const long long threads = N * 32;
    nseg_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(planeA, planeB, wpp, w_begin, w_end, sire, dam, N, nseg);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// VarA = 1/4 (Q_sire + Q_dam); VarD = 1/16 (P_sire + P_dam + 2 X); VarBV likewise from asub.
__global__ void assemble_kernel(const AssembleParams p) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.N * p.npairs) return;
    const long long x = idx / p.npairs;
    const int pr = (int)(idx - x * p.npairs);
    const int TT = p.T * p.T;
    const int i = p.pair_t1[pr] * p.T + p.pair_t2[pr];
    const long long ss = (long long)p.sire[x] * TT + i, ds = (long long)p.dam[x] * TT + i;
    double* o = p.out + idx * p.ncomp;
    o[0] = 0.25 * (p.QA[ss] + p.QA[ds]);
    if (p.ncomp >= 2) o[1] = 0.0625 * (p.PD[ss] + p.PD[ds] + 2.0 * p.X[x * TT + i]);
    if (p.ncomp >= 3) o[2] = 0.25 * (p.QB[ss] + p.QB[ds]);
}
cudaError_t launch_assemble(const AssembleParams& p, cudaStream_t st) {
    const long long total = p.N * p.npairs;
    if (total == 0) return cudaSuccess;
    assemble_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// predictCrosses() tidy output per cross (R/gmsFunctions.R:764-871), fused: predOf in {BV [, TGV]} x Trait in {[SELIND,] traits}
// -> (predMean, predVar, predSD, predUsefulness). VarBV = VarA (A, AD) or the asub component (DirDom); VarTGV = VarBV + VarDD
// (AD, :777-787) or VarA + VarD (DirDom, :788-799); MeanTGV = MeanBV for AD (:737-741); SELIND: w'Gw with G symmetrised from
// the upper triangle (:835-843) and w'mean (:812-822); usefulness = mean + stdSelInt * sd (:855-864).
__global__ void tidy_kernel(const TidyParams p) {
    const int nof = p.model == 0 ? 1 : 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.N * nof) return;
    const long long x = idx / nof;
    const int of = (int)(idx - x * nof);
    const int ntr = p.T + (p.w ? 1 : 0);
    const double* o = p.out + x * (long long)p.npairs * p.ncomp;
    const double* gs = p.gebv + (long long)p.sire[x] * p.T;
    const double* gd = p.gebv + (long long)p.dam[x] * p.T;
    const int bv = p.model == 2 ? 2 : 0;
    auto var_of = [&](int pr) { return of == 0 ? o[pr * p.ncomp + bv] : o[pr * p.ncomp] + o[pr * p.ncomp + 1]; };
    auto mean_of = [&](int t) {
        return (of == 1 && p.model == 2) ? p.mean_tgv[x * p.T + t] : (gs[t] + gd[t]) / 2.0;     // R/predCrossVar.R:369
    };
    double* dst = p.tidy + (idx * ntr) * 4;
    if (p.w) {
        double v = 0.0, mu = 0.0;
        for (int pr = 0; pr < p.npairs; ++pr) {
            const int a = p.pair_t1[pr], b = p.pair_t2[pr];
            v += (a == b ? 1.0 : 2.0) * p.w[a] * p.w[b] * var_of(pr);
        }
        for (int t = 0; t < p.T; ++t) mu += p.w[t] * mean_of(t);
        const double sd = sqrt(v);
        dst[0] = mu; dst[1] = v; dst[2] = sd; dst[3] = mu + p.std_sel_int * sd;
        dst += 4;
    }
    for (int t = 0; t < p.T; ++t, dst += 4) {
        const double v = var_of(t), mu = mean_of(t), sd = sqrt(v);
        dst[0] = mu; dst[1] = v; dst[2] = sd; dst[3] = mu + p.std_sel_int * sd;
    }
}
cudaError_t launch_tidy(const TidyParams& p, cudaStream_t st) {
    const long long total = p.N * (p.model == 0 ? 1 : 2);
    if (total == 0) return cudaSuccess;
    tidy_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// predCrossMeans (R/predCrossVar.R:336-399).
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
    return r;   // valid on thread 0
}

__global__ void gebv_kernel(const unsigned* __restrict__ planeA, const unsigned* __restrict__ planeB, long long wpp,
                            const int* __restrict__ plist, const double* __restrict__ emat, long long mp_total, int T,
                            double* __restrict__ gebv) {
    __shared__ double sh[8];
    const int par = plist ? plist[blockIdx.x] : (int)blockIdx.x;
    const unsigned* A = planeA + (long long)par * wpp;
    const unsigned* B = planeB + (long long)par * wpp;
    for (int t = 0; t < T; ++t) {
        const double* e = emat + (long long)t * mp_total;
        double s = 0.0;
        for (long long j = threadIdx.x; j < mp_total; j += blockDim.x) {
            const int dose = ((A[j >> 5] >> (j & 31)) & 1u) + ((B[j >> 5] >> (j & 31)) & 1u);
            s += dose * e[j];
        }
        const double tot = block_sum(s, sh);
        if (threadIdx.x == 0) gebv[(long long)blockIdx.x * T + t] = tot;
    }
}
cudaError_t launch_gebv(const unsigned* planeA, const unsigned* planeB, long long wpp, const int* plist,
                        long long nP, const double* emat, long long mp_total, int T, double* gebv, cudaStream_t st) {
    if (nP == 0) return cudaSuccess;
    gebv_kernel<<<(unsigned)nP, 256, 0, st>>>(planeA, planeB, wpp, plist, emat, mp_total, T, gebv);
    return cudaGetLastError();
}

__global__ void tgv_mean_kernel(const unsigned* __restrict__ planeA, const unsigned* __restrict__ planeB,
                                long long wpp, const int* __restrict__ sire, const int* __restrict__ dam,
                                const double* __restrict__ ea, const double* __restrict__ ed, long long mp_total,
                                int T, double* __restrict__ out) {
    __shared__ double sh[8];
    const long long x = blockIdx.x;
    const unsigned* As = planeA + (long long)sire[x] * wpp;
    const unsigned* Bs = planeB + (long long)sire[x] * wpp;
    const unsigned* Ad = planeA + (long long)dam[x] * wpp;
    const unsigned* Bd = planeB + (long long)dam[x] * wpp;
    for (int t = 0; t < T; ++t) {
        const double* a = ea + (long long)t * mp_total;
        const double* d = ed + (long long)t * mp_total;
        double s = 0.0;
        for (long long j = threadIdx.x; j < mp_total; j += blockDim.x) {
            const int sh_ = (int)(j & 31);
            const long long w = j >> 5;
            const double p1 = 0.5 * (((As[w] >> sh_) & 1u) + ((Bs[w] >> sh_) & 1u));
            const double p2 = 0.5 * (((Ad[w] >> sh_) & 1u) + ((Bd[w] >> sh_) & 1u));
            const double q = 1.0 - p1, y = p1 - p2;
            s += a[j] * (p1 - q - y) + d[j] * ((2.0 * p1 * q) + y * (p1 - q));
        }
        const double tot = block_sum(s, sh);
        if (threadIdx.x == 0) out[x * T + t] = tot;
    }
}
cudaError_t launch_tgv_mean(const unsigned* planeA, const unsigned* planeB, long long wpp, const int* sire,
                            const int* dam, long long N, const double* emat_add, const double* emat_dom,
                            long long mp_total, int T, double* mean_tgv, cudaStream_t st) {
    if (N == 0) return cudaSuccess;
    tgv_mean_kernel<<<(unsigned)N, 256, 0, st>>>(planeA, planeB, wpp, sire, dam, emat_add, emat_dom, mp_total, T, mean_tgv);
    return cudaGetLastError();
}

}  // namespace gms
