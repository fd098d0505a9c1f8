// The dominant kernel: fused "generate W -> FP64 tensor contraction -> quadratic-form reduce".
//
// For a tile of NU units (crosses, or parents) and T traits it evaluates, for every unit u and
// trait pair (t,s),
//        Z_u[t,s] = sum_j  e_t[j] mask_u[j] * ( sum_k L'[j,k] * e_s[k] mask_u[k] )
// where L' is the lower block triangle (diagonal block halved) of R or RoR for one chromosome,
// e_t the SNP effects and mask_u in {-1,0,+1}^m comes straight from the bit-packed haplotypes:
//   MASK_DELTA : delta_P = HapA - HapB            -> additive term    a' (R o dd') a
//   MASK_HET   : het_P   = HapA xor HapB          -> dominance, per-parent part
//   MASK_XI    : xi = delta_sire * delta_dam      -> dominance, per-cross part
// The symmetric result is x' M y = Z[t,s] + Z[s,t] (finalize_sym in assemble.cu).
// This replaces calcGameticLD + calcCrossLD + D*D + quadform of the reference
// (R/predCrossVar.R:209-222,261-282; R/helpers.R:302) without ever forming an m x m LD matrix.
//
// Structure (one CTA = one unit-tile x one range of row-tiles, 10 warps):
//   warps 8-9  producers: one thread issues cp.async.bulk (TMA engine, mbarrier complete_tx) for the
//              pre-tiled 32 KiB A chunk; both warps generate the 32 x 80 W chunk on the fly from
//              one haplotype word per unit and write it to shared memory in B-fragment order.
//   warps 0-7  consumers: mma.sync.m16n8k4.f64 (DMMA) on a 32 x 40 warp tile, accumulators in
//              registers; after each row-tile they fold the accumulators with e_t[j] mask_u[j]
//              (warp shuffles) into a per-warp T x 40 table in shared memory.
//   full/empty mbarrier ring of `stages` slots; no CTA-wide barrier in the main loop.
#include "common.cuh"

namespace gms {

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void dmma_16x8x4(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a0), "d"(a1), "d"(b0));
}

// nz / neg words of a unit for one 32-SNP word.
__device__ __forceinline__ void unit_mask(const ContractParams& p, int pa, int pb, long long w,
                                          unsigned& nz, unsigned& ng) {
    nz = 0u; ng = 0u;
    if (pa < 0) return;
    const unsigned A = __ldg(p.planeA + (long long)pa * p.wpp + w);
    const unsigned B = __ldg(p.planeB + (long long)pa * p.wpp + w);
    if (p.mode == MASK_XI) {
        const unsigned A2 = __ldg(p.planeA + (long long)pb * p.wpp + w);
        const unsigned B2 = __ldg(p.planeB + (long long)pb * p.wpp + w);
        nz = (A ^ B) & (A2 ^ B2);
        ng = ((~A & B) ^ (~A2 & B2)) & nz;
    } else {
        nz = A ^ B;
        ng = (p.mode == MASK_DELTA) ? (~A & B) : 0u;
    }
}

// non-blocking probe of an mbarrier phase (result consumed several hundred cycles later)
__device__ __forceinline__ unsigned mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}

struct Frag {
    double2 a0, a1;
    double b[5];
};

// One k4-step worth of operands for this warp: two 16-row A blocks (mb = wm and wm + 4) and five
// 8-column B blocks. A0/A1: whether the block has any real (non-padding) row in this row-tile.
template <bool A0, bool A1>
__device__ __forceinline__ void load_frag(Frag& f, const double* a, const double* b, int q) {
    if (A0) f.a0 = *reinterpret_cast<const double2*>(a + (q * 8) * 64);
    if (A1) f.a1 = *reinterpret_cast<const double2*>(a + (q * 8 + 4) * 64);
#pragma unroll
    for (int ni = 0; ni < 5; ++ni) f.b[ni] = b[(q * 10 + ni) * 32];
}
template <bool A0, bool A1>
__device__ __forceinline__ void mma_step(double (&acc)[2][5][4], const Frag& f) {
#pragma unroll
    for (int ni = 0; ni < 5; ++ni) {
        if (A0) dmma_16x8x4(acc[0][ni], f.a0.x, f.a0.y, f.b[ni]);
        if (A1) dmma_16x8x4(acc[1][ni], f.a1.x, f.a1.y, f.b[ni]);
    }
}

// Main loop of one row-tile for one consumer warp, software-pipelined across k4-steps AND across
// pipeline stages: the next stage's barrier is probed one whole stage early and its first operands
// are fetched before the last DMMAs of the current stage issue, so the FP64 tensor pipe never
// drains at a stage boundary.
template <bool A0, bool A1>
__device__ __forceinline__ void rowtile_mainloop(double (&acc)[2][5][4], const double* As, const double* Bs,
                                                 unsigned full0, unsigned empty0, int S, int nkc, unsigned& it,
                                                 int wm, int wn, int lane) {
    unsigned slot = it % S, ph = (it / S) & 1u;
    mbar_wait(full0 + 8 * slot, ph);
    const int aoff = (wm * 32 + lane) * 2, boff = (5 * wn) * 32 + lane;
    Frag cur, nxt;
    load_frag<A0, A1>(cur, As + (size_t)slot * CHUNK + aoff, Bs + (size_t)slot * BCHUNK + boff, 0);
    for (int kc = 0; kc < nkc; ++kc) {
        const unsigned nslot = (it + 1) % S, nph = ((it + 1) / S) & 1u;
        const bool has_next = kc + 1 < nkc;
        unsigned ok = 1u;
        if (has_next) ok = mbar_try(full0 + 8 * nslot, nph);
        const double* a = As + (size_t)slot * CHUNK + aoff;
        const double* b = Bs + (size_t)slot * BCHUNK + boff;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            load_frag<A0, A1>(nxt, a, b, q + 1);
            mma_step<A0, A1>(acc, cur);
            cur = nxt;
        }
        if (has_next) {
            if (!ok) mbar_wait(full0 + 8 * nslot, nph);
            load_frag<A0, A1>(nxt, As + (size_t)nslot * CHUNK + aoff, Bs + (size_t)nslot * BCHUNK + boff, 0);
        }
        mma_step<A0, A1>(acc, cur);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * slot);   // every read of this slot has completed
        cur = nxt;
        slot = nslot;
        ++it;
    }
}

// real (non-padding) extent of a row-tile: 16-row blocks with a real row, and k-chunks with a real k
__device__ __forceinline__ int rowtile_nact(const ChromDesc& cd, int I) {
    const int rows = cd.m - BM * I;
    return rows >= BM ? 8 : (rows + 15) >> 4;
}
__device__ __forceinline__ int rowtile_nkc(const ChromDesc& cd, int I) {
    const int full = 4 * (I + 1), real = (cd.m + BK - 1) / BK;
    return full < real ? full : real;
}

__global__ void __launch_bounds__(NTHREADS, 1) contract_kernel(const ContractParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = p.stages;
    const int T = p.T;
    const int NU = p.NU;
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + (size_t)S * CHUNK;
    double* Zacc = Bs + (size_t)S * BCHUNK;                       // [8 warps][40][T]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Zacc + MATH_WARPS * 40 * T);
    unsigned* emask = reinterpret_cast<unsigned*>(bars + 2 * S);  // [S + 1 buffers][NU][4 words][nz, ng]: the producers run at
                                                                  // most S chunks = S row-tiles ahead of the consumers' epilogue
    const unsigned full0 = smem_u32(bars);
    const unsigned empty0 = smem_u32(bars + S);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int split = blockIdx.y;
    // units: p.n_units of them, or - for the FP64 re-evaluation of the units that failed the int8 error bound (verify.cu) -
    // a count that only the device knows; then the grid is fixed and every CTA walks the tiles it owns
    const long long n_live = p.n_units_dev ? (long long)min(*p.n_units_dev, (int)p.n_units) : p.n_units;
    if ((long long)blockIdx.x * NU >= n_live) return;
    const int rt_begin = p.split_begin[split], rt_end = p.split_begin[split + 1];

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1 + PROD_WARPS);   // expect_tx arrive + one arrive per producer warp
            mbar_init(empty0 + 8 * s, MATH_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= MATH_WARPS) {
        // ------------------------------------------------------------------ producers
        const int pw = warp - MATH_WARPS;
        int colu[5], colt[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const int n = 40 * pw + 8 * b + g;
            if (n < NU * T) { colu[b] = n / T; colt[b] = n - colu[b] * T; }
            else { colu[b] = -1; colt[b] = 0; }
        }
        const int ulo = (40 * pw) / T;
        unsigned it = 0, rtc = 0;
        for (long long u0 = (long long)blockIdx.x * NU; u0 < n_live; u0 += (long long)gridDim.x * NU) {
        // lane keeps the parents of units ulo+lane and ulo+32+lane of this tile
        int pa0 = -1, pb0 = -1, pa1 = -1, pb1 = -1;
        {
            const int ua = ulo + lane, ub = ulo + 32 + lane;
            if (ua < NU && u0 + ua < n_live) { pa0 = p.unit_a[u0 + ua]; pb0 = (p.mode == MASK_XI) ? p.unit_b[u0 + ua] : 0; }
            if (ub < NU && u0 + ub < n_live) { pa1 = p.unit_a[u0 + ub]; pb1 = (p.mode == MASK_XI) ? p.unit_b[u0 + ub] : 0; }
        }
        for (int rt = rt_begin; rt < rt_end; ++rt, ++rtc) {
            const RowTile tile = p.rowtiles[rt];
            const ChromDesc cd = p.chrom[tile.c];
            const int nkc = rowtile_nkc(cd, tile.I);
            const double* src = p.tiles + cd.tile_off + chunks_before_rowtile(tile.I) * CHUNK;
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                const unsigned slot = it % S, ph = (it / S) & 1u;
                mbar_wait(empty0 + 8 * slot, ph ^ 1u);
                if (pw == 0 && lane == 0) {
                    mbar_arrive_expect_tx(full0 + 8 * slot, CHUNK * 8);
                    bulk_g2s(smem_u32(As + (size_t)slot * CHUNK), src + (size_t)kc * CHUNK, CHUNK * 8, full0 + 8 * slot);
                }
                if (kc == 0) {
                    // masks of this row-tile's 128 rows for the consumers' epilogue (published by the
                    // arrive on this stage's full barrier; double-buffered by row-tile parity)
                    unsigned* em = emask + (rtc % (unsigned)(S + 1)) * (NU * 8);
                    const long long wbase = (cd.pad_off >> 5) + 4 * tile.I;
                    for (int idx = pw * 32 + lane; idx < NU * 4; idx += PROD_WARPS * 32) {
                        const int u = idx >> 2, w4 = idx & 3;
                        int pa = -1, pb = 0;
                        if (u0 + u < n_live) { pa = p.unit_a[u0 + u]; pb = (p.mode == MASK_XI) ? p.unit_b[u0 + u] : 0; }
                        unsigned nz, ng;
                        unit_mask(p, pa, pb, wbase + w4, nz, ng);
                        em[idx * 2] = nz;
                        em[idx * 2 + 1] = ng;
                    }
                }
                const long long w = (cd.pad_off >> 5) + kc;
                unsigned nz0, ng0, nz1, ng1;
                unit_mask(p, pa0, pb0, w, nz0, ng0);
                unit_mask(p, pa1, pb1, w, nz1, ng1);
                const double* ebase = p.emat + cd.pad_off + 32 * kc;
                double* bdst = Bs + (size_t)slot * BCHUNK + (5 * pw) * 32 + lane;
#pragma unroll
                for (int b = 0; b < 5; ++b) {
                    const int rel = (colu[b] < 0 ? 0 : colu[b] - ulo);
                    const int srcl = rel & 31;
                    const unsigned a0 = __shfl_sync(0xffffffffu, nz0, srcl), a1 = __shfl_sync(0xffffffffu, nz1, srcl);
                    const unsigned c0 = __shfl_sync(0xffffffffu, ng0, srcl), c1 = __shfl_sync(0xffffffffu, ng1, srcl);
                    unsigned nz = (rel >> 5) ? a1 : a0, ng = (rel >> 5) ? c1 : c0;
                    if (colu[b] < 0) nz = 0u;
                    // integer-only select / sign flip: the FP64 pipe belongs to the DMMA warps
                    const unsigned long long* ecol =
                        reinterpret_cast<const unsigned long long*>(ebase + (long long)colt[b] * p.mp_total);
                    unsigned long long* bq = reinterpret_cast<unsigned long long*>(bdst);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k = 4 * q + tig;
                        unsigned long long e = 0ull;
                        if ((nz >> k) & 1u) e = __ldg(ecol + k);
                        e ^= (unsigned long long)((ng >> k) & 1u) << 63;
                        bq[(q * 10 + b) * 32] = e;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * slot);
            }
        }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        // warp (wm, wn): rows of the 16-row blocks mb = wm and wm + 4 (interleaved so that the padding
        // blocks of a chromosome's last row-tile are spread over the 4 SM sub-partitions), columns 40 wn ..
        const int wm = warp & 3, wn = warp >> 2;
        double acc[2][5][4];
        unsigned it = 0, rtc = 0;
        for (long long u0 = (long long)blockIdx.x * NU; u0 < n_live; u0 += (long long)gridDim.x * NU) {
        for (int i = threadIdx.x; i < MATH_WARPS * 40 * T; i += MATH_WARPS * 32) Zacc[i] = 0.0;
        asm volatile("bar.sync 1, %0;" ::"n"(MATH_WARPS * 32) : "memory");
        for (int rt = rt_begin; rt < rt_end; ++rt, ++rtc) {
            const RowTile tile = p.rowtiles[rt];
            const ChromDesc cd = p.chrom[tile.c];
            const int nkc = rowtile_nkc(cd, tile.I);
            const int nact = rowtile_nact(cd, tile.I);
            const bool act0 = wm < nact, act1 = wm + 4 < nact;
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 5; ++ni)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[mi][ni][r] = 0.0;

            if (act1) {
                rowtile_mainloop<true, true>(acc, As, Bs, full0, empty0, S, nkc, it, wm, wn, lane);
            } else if (act0) {
                rowtile_mainloop<true, false>(acc, As, Bs, full0, empty0, S, nkc, it, wm, wn, lane);
            } else {
                for (int kc = 0; kc < nkc; ++kc, ++it) {      // nothing to compute: keep the ring moving
                    const unsigned slot = it % S, ph = (it / S) & 1u;
                    mbar_wait(full0 + 8 * slot, ph);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8 * slot);
                }
                continue;
            }

            // ---- row-tile epilogue: Zacc[col][t] += sum_rows e_t[row] * mask_u(col)[row] * acc[row][col]
            const unsigned* em = emask + (rtc % (unsigned)(S + 1)) * (NU * 8);
            const int bit0 = 16 * (wm & 1) + g;            // row 16 (wm + 4 mi) + 8 h + g -> word (wm>>1) + 2 mi, bit bit0 + 8 h
#pragma unroll
            for (int ni = 0; ni < 5; ++ni) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int n = 40 * wn + 8 * ni + 2 * tig + j;
                    const bool live = n < NU * T;
                    const int u = live ? n / T : 0;
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        const uint2 mk = *reinterpret_cast<const uint2*>(em + (u * 4 + (wm >> 1) + 2 * mi) * 2);
                        const unsigned nz = live ? mk.x : 0u, ng = mk.y;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int bit = bit0 + 8 * h;
                            long long v = __double_as_longlong(acc[mi][ni][2 * h + j]);
                            if (!((nz >> bit) & 1u)) v = 0ll;
                            v ^= (long long)((ng >> bit) & 1u) << 63;
                            acc[mi][ni][2 * h + j] = __longlong_as_double(v);
                        }
                    }
                }
            }
            const double* erow = p.emat + cd.pad_off + BM * tile.I + 16 * wm + g;   // + 8 h + 64 mi
            double* zw = Zacc + (size_t)(warp * 40) * T;
            double e00 = __ldg(erow), e01 = __ldg(erow + 8), e10 = __ldg(erow + 64), e11 = __ldg(erow + 72);
            for (int t = 0; t < T; ++t) {
                const double c00 = e00, c01 = e01, c10 = e10, c11 = e11;
                if (t + 1 < T) {                               // prefetch the next trait's effects
                    const double* et = erow + (long long)(t + 1) * p.mp_total;
                    e00 = __ldg(et); e01 = __ldg(et + 8); e10 = __ldg(et + 64); e11 = __ldg(et + 72);
                }
#pragma unroll
                for (int ni = 0; ni < 5; ++ni) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        double s = c00 * acc[0][ni][j];
                        s = fma(c01, acc[0][ni][2 + j], s);
                        s = fma(c10, acc[1][ni][j], s);
                        s = fma(c11, acc[1][ni][2 + j], s);
                        s += __shfl_xor_sync(0xffffffffu, s, 4);
                        s += __shfl_xor_sync(0xffffffffu, s, 8);
                        s += __shfl_xor_sync(0xffffffffu, s, 16);
                        if (g == 0) zw[(8 * ni + 2 * tig + j) * T + t] += s;
                    }
                }
            }
        }

        // ---- write Z for this (tile, split): fixed summation order over the 4 M-warps
        asm volatile("bar.sync 1, %0;" ::"n"(MATH_WARPS * 32) : "memory");
        const int TT = T * T;
        for (int idx = threadIdx.x; idx < NU * TT; idx += MATH_WARPS * 32) {
            const int u = idx / TT;
            const int rem = idx - u * TT;
            const int t = rem / T, s = rem - t * T;
            const int n = u * T + s;
            const int wn2 = n / 40, cl = n - wn2 * 40;
            double v = 0.0;
#pragma unroll
            for (int m4 = 0; m4 < 4; ++m4) v += Zacc[(size_t)((wn2 * 4 + m4) * 40 + cl) * T + t];
            const long long unit = u0 + u;
            if (unit < n_live) p.Z[((long long)split * p.n_units + unit) * TT + rem] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(MATH_WARPS * 32) : "memory");     // Zacc is re-zeroed for the CTA's next tile
        }
    }
}

size_t contract_smem_bytes(int T, int stages) {
    const int NU = BN / T;
    return (size_t)stages * (CHUNK + BCHUNK) * 8 + (size_t)MATH_WARPS * 40 * T * 8 + (size_t)2 * stages * 8 +
           (size_t)(stages + 1) * NU * 8 * 4;
}

int contract_pick_stages(int T, size_t smem_limit) {
    for (int s = 4; s >= 2; --s)
        if (contract_smem_bytes(T, s) <= smem_limit) return s;
    return 0;
}

cudaError_t launch_contract(const ContractParams& p, int ntiles, int nsplit, cudaStream_t st) {
    const size_t smem = contract_smem_bytes(p.T, p.stages);
    cudaError_t e = cudaFuncSetAttribute(contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(ntiles, nsplit, 1);
    contract_kernel<<<grid, NTHREADS, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace gms
