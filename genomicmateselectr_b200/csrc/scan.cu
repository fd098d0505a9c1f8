// Map-structured fast path (optional, behind the dense contraction).
//
// When recombFreqMat comes from a genetic map (the only way the reference's users make it:
// 1-2*genmap2recombfreq(map), R/helpers.R:204-213), within a chromosome with sorted positions x
//      R[j,k]   = exp(-0.02 |x_j - x_k|) = a_j b_k  (k <= j),  a_j = exp(-c x_j), b_k = exp(c x_k), c = 0.02
//      RoR[j,k] = exp(-0.04 |x_j - x_k|)                                                    c = 0.04
// so every quadratic form is a prefix-sum recurrence instead of an m x m contraction:
//      w' M v = sum_j [ w_j v_j + a_j w_j G_j(v) + a_j v_j G_j(w) ],   G_j(v) = sum_{k<j} b_k v_k
// and only the SNPs where the unit's mask is non-zero (delta, het or xi) contribute.  Cost per unit:
// O(nnz * T^2) instead of O(m_c^2 * T): 10^2 - 10^3 x fewer FP64 operations than the dense path.
// One thread = one unit x one range of chromosomes; a TB x TB block of trait pairs lives in registers.
#include "kernels.cuh"

namespace gms {

// S[j][3][T]: e_t[j], a_j e_t[j], b_j e_t[j] on the padded SNP axis
__global__ void scan_prepare_kernel(const double* __restrict__ emat, const double* __restrict__ apos,
                                    const double* __restrict__ bpos, long long mp_total, int T,
                                    double* __restrict__ S) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mp_total) return;
    const double a = apos[j], b = bpos[j];
    for (int t = 0; t < T; ++t) {
        const double e = emat[(long long)t * mp_total + j];
        S[(j * 3 + 0) * T + t] = e;
        S[(j * 3 + 1) * T + t] = a * e;
        S[(j * 3 + 2) * T + t] = b * e;
    }
}
cudaError_t launch_scan_prepare(const double* emat, const double* apos, const double* bpos, long long mp_total, int T,
                                double* S, cudaStream_t st) {
    scan_prepare_kernel<<<(unsigned)((mp_total + 255) / 256), 256, 0, st>>>(emat, apos, bpos, mp_total, T, S);
    return cudaGetLastError();
}

template <int TB, bool SINGLE>
__global__ void __launch_bounds__(128) scan_kernel(const ScanParams p) {
    const long long unit = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (unit >= p.n_units) return;
    const int split = blockIdx.y;
    // trait block pair (A <= B) for T > TB
    const int nblk = (p.T + TB - 1) / TB;
    int ba = 0, bb = SINGLE ? 0 : blockIdx.z;
    if (!SINGLE) {
        while (bb >= nblk - ba) { bb -= nblk - ba; ++ba; }
        bb += ba;
    }
    const int ta0 = ba * TB, tb0 = bb * TB;
    const bool diag = SINGLE || ba == bb;      // SINGLE (T <= 6): one diagonal block, upper triangle only (compile time)
    const int nA = min(TB, p.T - ta0), nB = min(TB, p.T - tb0);
    const int T = p.T;

    const int pa = p.unit_a[unit];
    const int pb = (p.mode == MASK_XI) ? p.unit_b[unit] : 0;
    const unsigned* A1 = p.planeA + (long long)pa * p.wpp;
    const unsigned* B1 = p.planeB + (long long)pa * p.wpp;
    const unsigned* A2 = p.planeA + (long long)pb * p.wpp;
    const unsigned* B2 = p.planeB + (long long)pb * p.wpp;

    double X[TB][TB];
#pragma unroll
    for (int t = 0; t < TB; ++t)
#pragma unroll
        for (int s = 0; s < TB; ++s) X[t][s] = 0.0;

    for (int c = p.split_begin[split]; c < p.split_begin[split + 1]; ++c) {
        const ChromDesc cd = p.chrom[c];
        double GA[TB], GB[TB];
#pragma unroll
        for (int t = 0; t < TB; ++t) { GA[t] = 0.0; GB[t] = 0.0; }
        const int w0 = cd.pad_off >> 5, nw = (cd.m + 31) >> 5;
        for (int wq = 0; wq < nw; wq += 4) {           // chromosomes start on 128-SNP boundaries: 16-byte aligned
            const uint4 a1 = *reinterpret_cast<const uint4*>(A1 + w0 + wq);
            const uint4 b1 = *reinterpret_cast<const uint4*>(B1 + w0 + wq);
            uint4 nzv, ngv;
            if (p.mode == MASK_XI) {
                const uint4 a2 = *reinterpret_cast<const uint4*>(A2 + w0 + wq);
                const uint4 b2 = *reinterpret_cast<const uint4*>(B2 + w0 + wq);
                nzv.x = (a1.x ^ b1.x) & (a2.x ^ b2.x); ngv.x = ((~a1.x & b1.x) ^ (~a2.x & b2.x)) & nzv.x;
                nzv.y = (a1.y ^ b1.y) & (a2.y ^ b2.y); ngv.y = ((~a1.y & b1.y) ^ (~a2.y & b2.y)) & nzv.y;
                nzv.z = (a1.z ^ b1.z) & (a2.z ^ b2.z); ngv.z = ((~a1.z & b1.z) ^ (~a2.z & b2.z)) & nzv.z;
                nzv.w = (a1.w ^ b1.w) & (a2.w ^ b2.w); ngv.w = ((~a1.w & b1.w) ^ (~a2.w & b2.w)) & nzv.w;
            } else {
                nzv.x = a1.x ^ b1.x; nzv.y = a1.y ^ b1.y; nzv.z = a1.z ^ b1.z; nzv.w = a1.w ^ b1.w;
                if (p.mode == MASK_DELTA) { ngv.x = ~a1.x & b1.x; ngv.y = ~a1.y & b1.y; ngv.z = ~a1.z & b1.z; ngv.w = ~a1.w & b1.w; }
                else { ngv.x = ngv.y = ngv.z = ngv.w = 0u; }
            }
            const unsigned nzs[4] = {nzv.x, nzv.y, nzv.z, nzv.w};
            const unsigned ngs[4] = {ngv.x, ngv.y, ngv.z, ngv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned nz = nzs[i];
                const unsigned ng = ngs[i];
                while (nz) {
                    const int bit = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const long long j = ((long long)(w0 + wq + i) << 5) + bit;
                    const long long sgn = (long long)((ng >> bit) & 1u) << 63;
                    const double* s = p.S + j * 3 * T;
                    double eA[TB], aA[TB], bA[TB], eB[TB], aB[TB], bB[TB];
#pragma unroll
                    for (int t = 0; t < TB; ++t) {
                        if (t < nA) {
                            eA[t] = __ldg(s + ta0 + t);
                            aA[t] = __longlong_as_double(__double_as_longlong(__ldg(s + T + ta0 + t)) ^ sgn);
                            bA[t] = __longlong_as_double(__double_as_longlong(__ldg(s + 2 * T + ta0 + t)) ^ sgn);
                        } else { eA[t] = aA[t] = bA[t] = 0.0; }
                    }
                    if (diag) {
#pragma unroll
                        for (int t = 0; t < TB; ++t) { eB[t] = eA[t]; aB[t] = aA[t]; bB[t] = bA[t]; }
                    } else {
#pragma unroll
                        for (int t = 0; t < TB; ++t) {
                            if (t < nB) {
                                eB[t] = __ldg(s + tb0 + t);
                                aB[t] = __longlong_as_double(__double_as_longlong(__ldg(s + T + tb0 + t)) ^ sgn);
                                bB[t] = __longlong_as_double(__double_as_longlong(__ldg(s + 2 * T + tb0 + t)) ^ sgn);
                            } else { eB[t] = aB[t] = bB[t] = 0.0; }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < TB; ++t)
#pragma unroll
                        for (int u = 0; u < TB; ++u)
                            if (SINGLE ? (u >= t) : (!diag || u >= t))
                                X[t][u] += fma(aA[t], GB[u], fma(aB[u], GA[t], eA[t] * eB[u]));
#pragma unroll
                    for (int t = 0; t < TB; ++t) { GA[t] += bA[t]; GB[t] += bB[t]; }
                }
            }
        }
    }
    double* out = p.Xpart + ((long long)split * p.n_units + unit) * T * T;
#pragma unroll
    for (int t = 0; t < TB; ++t)
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            if (t < nA && u < nB && (SINGLE ? (u >= t) : (!diag || u >= t))) {
                out[(ta0 + t) * T + (tb0 + u)] = X[t][u];
                out[(tb0 + u) * T + (ta0 + t)] = X[t][u];
            }
        }
}

cudaError_t launch_scan(const ScanParams& p, int nsplit, cudaStream_t st) {
    if (p.n_units == 0) return cudaSuccess;
    const unsigned gx = (unsigned)((p.n_units + 127) / 128);
#define GMS_SCAN(TBV, SGL)                                                     \
    {                                                                          \
        const int nblk = (p.T + TBV - 1) / TBV;                                \
        dim3 grid(gx, nsplit, nblk * (nblk + 1) / 2);                          \
        scan_kernel<TBV, SGL><<<grid, 128, 0, st>>>(p);                        \
    }
    switch (p.T) {
        case 1: GMS_SCAN(1, true) break;
        case 2: GMS_SCAN(2, true) break;
        case 3: GMS_SCAN(3, true) break;
        case 4: GMS_SCAN(4, true) break;
        case 5: GMS_SCAN(5, true) break;
        case 6: GMS_SCAN(6, true) break;
        default: GMS_SCAN(4, false) break;     // T >= 7: 4 x 4 blocks of trait pairs, one grid.z slice per block pair
    }
#undef GMS_SCAN
    return cudaGetLastError();
}

}  // namespace gms
