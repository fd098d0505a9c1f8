// The accuracy contract of the int8 digit-plane path (GMS_MODE_INT8_TENSOR): an a-posteriori error bound per unit
// (parent or cross), evaluated on the device right after each table, and a list of the units that fail it. The
// failing units are re-evaluated by the FP64 contraction (contract.cu) in the same stream; see api.cu.
//
// Why a bound is needed. The int8 path rounds G_s[j,k] = L'[j,k] e_s[k] to I8_FRAC = 49 fractional bits RELATIVE TO THE ROW
// SCALE scale_s[j] = 2^E >= max_k |G_s[j,k]| (i8path.cu, i8digits.cuh; 47 bits in the six-plane build variant). That is fixed point, not floating point: a SNP with
// a huge effect sets the scale of every row it reaches, and a unit in which that SNP does not segregate (mask 0) sums only
// small terms that each carry the absolute error u * scale_s[j], u = I8_U = 2^-50 + 2^-52: 2^-50 from rounding to
// nearest (i8_scale_kernel picks E so that no digit ever saturates) plus the 53-bit truncation of a row sum in the epilogue.
// For unit x with mask mu in {-1,0,1}^m, row j sums over the non-zero mask entries k in the 64-blocks its row tile owns
// (about half of the chromosome's, common.cuh), at most n_c(x) = the unit's non-zero entries on the chromosome, so
//     |Z[t][s] - exact| <= u * sum_j |e_t[j] mu_j| * scale_s[j] * n_c  <=  u * Smax_s * B_t(x),
//     B_t(x) = sum_c n_c(x) sum_{j in c} |e_t[j] mu_j|,   Smax_s = largest row scale of trait s.
// The table entry is X[t,s] = Z[t][s] + Z[s][t], so a unit PASSES when
//     2 u * max_t (B_t / sqrt|X[t,t]|) * max_s (Smax_s / sqrt|X[s,s]|)  <=  tau              (tau = 2^-31 ~ 4.7e-10)
// which implies |error of X[t,s]| <= tau * sqrt(X[t,t] X[s,s]) for every trait pair: every variance to tau relative,
// every covariance to tau of its variance scale (the guard of tests/util.py:rel_err is 1e-6 of that scale at 1e-9).
// By Cauchy-Schwarz the same then holds for VarA = (Q_S + Q_D) / 4 and VarD = (P_S + P_D + 2 X) / 16. It is a worst-case
// bound (all rounding errors maximal and aligned); the error actually made is ~sqrt(n) times smaller. On N(0,1)
// effects the bound sits at 3.5 % (crosses) - 6.5 % (parents) of tau at cassava scale and at 7 % - 18 % at the large
// configuration (100k SNPs, T = 10): nothing fails.
// A unit also fails when it has >= 16384 non-zero mask entries on one chromosome: from 32768 on the exact 64-bit
// recombination of the digit sums (i8digits.cuh: i8_combine) could overflow; the limit keeps a factor of two in hand.
#include "kernels.cuh"

namespace gms {

constexpr int VER_THREADS = 128;     // one thread = one unit
constexpr int VER_TG = 8;            // traits per pass (accumulators in registers)
constexpr int VER_CH = 512;          // SNPs of |e_t| staged in shared memory per step (8 x 512 doubles = 32 KiB)

// Smax[s][c] = 2^E > max_j rowmax_s[j] over the rows of chromosome c (0 if every row is zero). Runs on the row-maximum
// buffer BEFORE i8_scale_kernel turns it into scales, with the same frexp rule. grid (C_shard, T).
__global__ void i8_smax_kernel(const double* __restrict__ rowmax, const ChromDesc* __restrict__ chrom, int c_begin, int C,
                               long long mp_total, double* __restrict__ smax) {
    const int c = c_begin + blockIdx.x, s = blockIdx.y;
    const ChromDesc cd = chrom[c];
    double mx = 0.0;
    for (int j = threadIdx.x; j < cd.m; j += blockDim.x) mx = fmax(mx, rowmax[(long long)s * mp_total + cd.pad_off + j]);
    __shared__ double sh[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, sh[i]);
        int e = 0;
        double v = 0.0;
        if (mx > 0x1p-900) { e = i8_scale_exp(mx); v = ldexp(1.0, e); }   // = i8_scale_kernel's rule
        smax[(long long)s * C + c] = v;
    }
}

// Step 1: B_t(x) restricted to one chromosome and the chromosome's non-zero count. grid (unit blocks, chromosomes of the
// shard): 18 x more CTAs and 18 x shorter serial walks than one thread per unit over the whole genome (measured: 2 ms ->
// ~0.15 ms per table at cassava scale); partial sums per chromosome, added in a fixed order by step 2 (no atomics).
// |e_t[j]| is staged TRAIT-CONTIGUOUS (se[j][0 .. TG)): a set mask bit costs TG / 2 128-bit shared-memory loads of one row
// instead of one 64-bit load per trait from TG different rows. The lanes of a warp sit at unrelated j, so every such load is
// a ~6-way bank conflict - the loads were the whole cost of this kernel (1.27 ms per 50k cassava crosses with eight 64-bit
// loads per bit). TG = T rounded up to even (at most VER_TG per pass); the sums run over j in the same order as before.
template <int TG>
__global__ void __launch_bounds__(VER_THREADS) i8_verify_accumulate_kernel(const VerifyParams p) {
    static_assert(TG % 2 == 0 && TG <= VER_TG, "trait group");
    __shared__ __align__(16) double se[VER_CH * TG];
    const long long unit = (long long)blockIdx.x * VER_THREADS + threadIdx.x;
    const int c = p.c_begin + blockIdx.y;
    const bool live = unit < p.n_units;
    int pa = 0, pb = 0;
    if (live) { pa = p.unit_a[unit]; pb = (p.mode == MASK_XI) ? p.unit_b[unit] : 0; }
    const uint4* A1 = reinterpret_cast<const uint4*>(p.planeA + (long long)pa * p.wpp);
    const uint4* B1 = reinterpret_cast<const uint4*>(p.planeB + (long long)pa * p.wpp);
    const uint4* A2 = reinterpret_cast<const uint4*>(p.planeA + (long long)pb * p.wpp);
    const uint4* B2 = reinterpret_cast<const uint4*>(p.planeB + (long long)pb * p.wpp);
    const ChromDesc cd = p.chrom[c];
    const bool xi = p.mode == MASK_XI;
    auto load_nz = [&](long long q, uint4& nz) {
        const uint4 a = __ldg(A1 + q), b = __ldg(B1 + q);
        nz = make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w);
        if (xi) {
            const uint4 a2 = __ldg(A2 + q), b2 = __ldg(B2 + q);
            nz.x &= a2.x ^ b2.x; nz.y &= a2.y ^ b2.y; nz.z &= a2.z ^ b2.z; nz.w &= a2.w ^ b2.w;
        }
    };
    double* vb = p.part_b + ((long long)blockIdx.y * p.n_units + unit) * p.T;
    for (int t0 = 0; t0 < p.T; t0 += TG) {
        const int ng = min(TG, p.T - t0);
        double acc[TG];
#pragma unroll
        for (int i = 0; i < TG; ++i) acc[i] = 0.0;
        int nnz = 0;                                           // non-zero mask entries of this chromosome so far
        const long long qbase = (long long)cd.pad_off >> 7;   // uint4 index (128 SNPs)
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (live) load_nz(qbase, nxt);
        for (int j0 = 0; j0 < cd.mp; j0 += VER_CH) {
            const int nj = min(VER_CH, cd.mp - j0);            // multiple of 128
            __syncthreads();
            for (int j = threadIdx.x; j < nj; j += VER_THREADS) {      // one row per thread: coalesced reads per trait, 128-bit stores
                double v[TG];
#pragma unroll
                for (int i = 0; i < TG; ++i)
                    v[i] = i < ng ? fabs(p.emat[(long long)(t0 + i) * p.mp_total + cd.pad_off + j0 + j]) : 0.0;
#pragma unroll
                for (int i = 0; i < TG; i += 2) *reinterpret_cast<double2*>(se + j * TG + i) = make_double2(v[i], v[i + 1]);
            }
            __syncthreads();
            if (!live) continue;
            for (int q = 0; q < nj / 128; ++q) {
                const uint4 cur = nxt;
                const long long qn = qbase + (j0 >> 7) + q + 1;
                if (qn < qbase + (cd.mp >> 7)) load_nz(qn, nxt);          // in flight while this group is processed
                const unsigned nzw[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    unsigned nz = nzw[w];
                    while (nz) {
                        const int j = q * 128 + w * 32 + (__ffs(nz) - 1);
                        nz &= nz - 1;
                        ++nnz;
                        const double2* row = reinterpret_cast<const double2*>(se + j * TG);
#pragma unroll
                        for (int i = 0; i < TG / 2; ++i) {
                            const double2 e = row[i];
                            acc[2 * i] += e.x;
                            acc[2 * i + 1] += e.y;
                        }
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int i = 0; i < TG; ++i)
                if (i < ng) vb[t0 + i] = acc[i] * (double)nnz;        // n_j <= the unit's non-zeros on the chromosome
            if (t0 == 0) p.part_n[(long long)blockIdx.y * p.n_units + unit] = nnz;
        }
    }
}

// Step 2: one thread per unit: sum the chromosomes in order, normalise by the unit's variances, decide.
__global__ void i8_verify_check_kernel(const VerifyParams p) {
    const long long unit = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (unit >= p.n_units) return;
    const int pa = p.unit_a[unit], pb = (p.mode == MASK_XI) ? p.unit_b[unit] : 0;
    const long long row = p.rows_by_parent ? (long long)pa : unit;
    const double* X = p.table + row * (long long)(p.T * p.T);
    const int nc = p.c_end - p.c_begin;
    double amax = 0.0, smax = 0.0;
    int nnz_max = 0;
    for (int c = 0; c < nc; ++c) nnz_max = max(nnz_max, p.part_n[(long long)c * p.n_units + unit]);
    for (int t = 0; t < p.T; ++t) {
        double b = 0.0, sm = 0.0;
        for (int c = 0; c < nc; ++c) {
            b += p.part_b[((long long)c * p.n_units + unit) * p.T + t];
            sm = fmax(sm, p.smax[(long long)t * p.C + p.c_begin + c]);
        }
        const double v = fabs(X[t * p.T + t]);
        const double r = v > 0.0 ? rsqrt(v) : 0.0;
        // zero variance: fine if the unit contributes nothing for this trait, otherwise fail (inf)
        amax = fmax(amax, b > 0.0 ? (v > 0.0 ? b * r : INFINITY) : 0.0);
        smax = fmax(smax, sm > 0.0 ? (v > 0.0 ? sm * r : INFINITY) : 0.0);
    }
    const double bound = 2.0 * I8_U * amax * smax;
    // NaN-safe: anything that is not provably within tau fails
    const bool ok = (bound <= p.tau || (amax == 0.0 || smax == 0.0)) && nnz_max < 16384;
    if (p.bound_out) p.bound_out[unit] = bound;
    if (!ok) {
        const int slot = atomicAdd(p.counter, 1);
        if (slot < p.cap) {
            p.list_a[slot] = pa;
            p.list_b[slot] = pb;
            p.list_row[slot] = (int)row;
        }
    }
}

cudaError_t launch_i8_smax(const double* rowmax, const ChromDesc* chrom_d, int c_begin, int c_end, int C, long long mp_total,
                           int T, double* smax, cudaStream_t st) {
    if (c_end <= c_begin) return cudaSuccess;
    i8_smax_kernel<<<dim3(c_end - c_begin, T), 256, 0, st>>>(rowmax, chrom_d, c_begin, C, mp_total, smax);
    return cudaGetLastError();
}

size_t i8_verify_scratch_doubles(long long n_units, int nchrom, int T) { return (size_t)n_units * nchrom * T; }

cudaError_t launch_i8_verify(const VerifyParams& p, cudaStream_t st) {
    if (p.n_units == 0) return cudaSuccess;
    const unsigned gx = (unsigned)((p.n_units + VER_THREADS - 1) / VER_THREADS);
    const dim3 grid(gx, p.c_end - p.c_begin);
    if (p.T <= 2) i8_verify_accumulate_kernel<2><<<grid, VER_THREADS, 0, st>>>(p);
    else if (p.T <= 4) i8_verify_accumulate_kernel<4><<<grid, VER_THREADS, 0, st>>>(p);
    else if (p.T <= 6) i8_verify_accumulate_kernel<6><<<grid, VER_THREADS, 0, st>>>(p);
    else i8_verify_accumulate_kernel<VER_TG><<<grid, VER_THREADS, 0, st>>>(p);
    i8_verify_check_kernel<<<(unsigned)((p.n_units + 127) / 128), 128, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace gms
