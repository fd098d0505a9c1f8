// Second-generation int8 tcgen05 contraction kernel (the default for the int8 path; i8path.cu documents the arithmetic).
//
// What changed against i8_contract_kernel and why. Measured on B200, 18,944 cassava crosses: 16.1 ms; MMAs alone 11.5 ms;
// epilogue alone 12.0 ms; bulk copies alone 7.8 ms; ideal tensor time 8.9 ms. The three phases were serialised per
// (tile, trait) unit because one 448-column accumulator fills TMEM; the epilogue spent its time in quarter-rate I2F.F64; the
// single-lane producer / MMA loops cost ~570 cycles of instructions per 448-cycle stage; and the xi tiles - which, unlike the
// digit tiles, no other SM reads at the same time - cost five times more per byte on the L2 -> SM path. Here
//   * the row tile is 32 SNP rows, so one accumulator is 7 x 32 = 224 TMEM columns and TWO fit: the MMAs of unit u + 1 run
//     while the epilogue drains unit u;
//   * 16 epilogue warps (four per TMEM lane quarter, 8 SNP rows each) instead of 8, int -> FP64 by exponent-bias bit tricks
//     (one LOP + one full-rate DADD) instead of I2F.F64, the 2^-49 digit weight folded into the staged row scale, the row
//     data and haplotype words of the next tile prefetched into registers while the current tile is processed;
//   * MC: the kernel is launched as clusters of two CTAs that share the SAME 128 crosses and take the two 32-row halves of
//     every 64-row block; each CTA fetches half of the xi bytes of a stage and multicasts them to both (the xi traffic per
//     MAC is that of a 64-row tile), its digit tile is its own;
//   * a pipeline stage is two 64-k blocks = three bulk copies and one commit (the digit image keeps the k-blocks of a
//     (row tile, trait) contiguous), and the producer / MMA warps walk their loops warp-uniformly with one elected lane
//     issuing.
// Layout of the digit image of one 64-row block row J of a chromosome: [trait s][row half h][kb <= J][224-row K-major
// operand], see i8_build_digits_kernel (layout 2).
#include "kernels.cuh"
#include "i8ptx.cuh"

namespace gms {

constexpr int V2_M = 128;                        // crosses per CTA (TMEM lanes)
constexpr int V2_ROWS = 32;                      // SNP rows j per accumulator
constexpr int V2_KB = 64;                        // k per pipeline stage
constexpr int V2_DIG = 7;
constexpr int V2_N = V2_DIG * V2_ROWS;           // 224 accumulator columns
constexpr int V2_A_BYTES = V2_M * V2_KB;         // 8 KiB
constexpr int V2_B_BYTES = V2_N * V2_KB;         // 14 KiB
constexpr int V2_BLK_BYTES = 2 * V2_B_BYTES;     // both row halves of a 64 x 64 block
constexpr int V2_KS = 2;                         // 64-k blocks per pipeline stage (one pair of bulk copies, one commit)
constexpr int V2_B_OFF = V2_KS * V2_A_BYTES;     // stage = [xi kb0][xi kb1][digits kb0][digits kb1]
constexpr int V2_STAGE_BYTES = V2_KS * (V2_A_BYTES + V2_B_BYTES);     // 44 KiB
constexpr int V2_STAGES = 4;                     // 176 KiB
constexpr int V2_MAX_STAGES = V2_STAGES;
constexpr int V2_EPI_WARPS = 16;
constexpr int V2_EPI_THREADS = 32 * V2_EPI_WARPS;
constexpr int V2_THREADS = 32 * (V2_EPI_WARPS + 2);
constexpr int V2_BUF_COLS = 256;                 // TMEM column stride between the two accumulators
constexpr int V2_NPART = 4;                      // partial slabs per split and CTA: one per 8-row column group (x 2 row halves with MC)

__device__ __forceinline__ void v2_tmem_ld4(unsigned taddr, int (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
// shared-memory accesses by 32-bit shared-window address (a generic pointer makes the compiler re-derive the window base)
__device__ __forceinline__ double2 v2_lds128(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void v2_sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

__device__ __forceinline__ void v2_bulk_g2s_mc(unsigned dst, const void* src, unsigned bytes, unsigned bar, unsigned short mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void v2_commit_mc(unsigned bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

// One 64-row block row (c, J) of the CTA's work list. The CTA copies its range of the list into shared memory once
// (V2_TTAB entries; the host plan keeps every split below that), so that no role chases jtiles -> chrom -> gbase through
// global memory on its per-unit critical path.
constexpr int V2_TTAB = 512;
struct V2Tile { int J, nh, nkb; long long prow0; long long gblk; };   // prow0: padded index of the block row's first SNP row;
                                                                       // gblk: index of its first (block, trait) digit image
__device__ __forceinline__ int4 v2_tile_pack(const I8Params& p, int jt) {
    const RowTile t = p.jtiles[jt];
    const ChromDesc cd = p.chrom[t.c];
    const int nh = (cd.m - 64 * t.I > V2_ROWS) ? 2 : 1;       // the second half of the last block may be all padding
    const int nkb = min(t.I + 1, (cd.m + V2_KB - 1) / V2_KB);
    const long long gblk = p.gbase[t.c] / V2_BLK_BYTES + (long long)(t.I * (t.I + 1) / 2) * p.T;
    return make_int4(t.I | (nh << 16) | (nkb << 18), cd.pad_off + 64 * t.I, (int)(gblk & 0xFFFFFFFFll), (int)(gblk >> 32));
}
__device__ __forceinline__ V2Tile v2_tile_unpack(const int4 e) {
    V2Tile r;
    r.J = e.x & 0xFFFF; r.nh = (e.x >> 16) & 3; r.nkb = e.x >> 18;
    r.prow0 = e.y;
    r.gblk = ((long long)e.w << 32) | (unsigned)e.z;
    return r;
}
// every role walks the units (32-row tile, trait) in the same order
// (hfix < 0: both row halves of every block, skipping an all-padding second half; hfix >= 0: that half only - the CTA's
// share in a row-half cluster, where both CTAs must walk identical k loops)
template <bool SOUTER, class F>
__device__ __forceinline__ void v2_for_units(const int4* ttab, int T, int jt_begin, int jt_end, int hfix, F f) {
    if (!SOUTER) {
        for (int jt = jt_begin; jt < jt_end; ++jt) {
            const V2Tile t = v2_tile_unpack(ttab[jt - jt_begin]);
            const int h0 = hfix < 0 ? 0 : hfix, h1 = hfix < 0 ? t.nh : hfix + 1;
            for (int h = h0; h < h1; ++h)
                for (int s = 0; s < T; ++s) f(t, h, s);
        }
    } else {
        for (int s = 0; s < T; ++s)
            for (int jt = jt_begin; jt < jt_end; ++jt) {
                const V2Tile t = v2_tile_unpack(ttab[jt - jt_begin]);
                const int h0 = hfix < 0 ? 0 : hfix, h1 = hfix < 0 ? t.nh : hfix + 1;
                for (int h = h0; h < h1; ++h) f(t, h, s);
            }
    }
}

// TT = T (SOUTER = false: tile outer, trait inner, the cross's T x T table in registers; T <= 5 at 112 registers)
// or TT >= T (SOUTER = true: trait outer, one column of the table in registers).
// IA (with SOUTER): integer accumulation - no floating-point instruction inside the kernel's main loop (see v2_combine).
template <int TT, bool SOUTER, bool MC, bool IA>
__global__ void __launch_bounds__(V2_THREADS, 1) i8v2_kernel(const I8Params p) {
    static_assert(!IA || SOUTER, "integer accumulation uses the trait-outer order");
    const int T = (SOUTER && !(IA && TT <= 6)) ? p.T : TT;       // IA is instantiated with TT == T up to 6
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int STAGES = V2_STAGES, STAGE_BYTES = V2_STAGE_BYTES;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * V2_MAX_STAGES + 5);
    int4* ttab = reinterpret_cast<int4*>(bars + 2 * V2_MAX_STAGES + 6);
    double* rowbuf = reinterpret_cast<double*>(ttab + V2_TTAB);
    const unsigned full0 = i8_smem_u32(bars), empty0 = i8_smem_u32(bars + V2_MAX_STAGES);
    const unsigned accfull0 = i8_smem_u32(bars + 2 * V2_MAX_STAGES), accempty0 = i8_smem_u32(bars + 2 * V2_MAX_STAGES + 2);
    const unsigned gobar = i8_smem_u32(bars + 2 * V2_MAX_STAGES + 4);       // FP64 variant: "the FP64 burst of the previous unit is over"
    constexpr bool GAP = SOUTER && !IA;       // see the trait-outer FP64 epilogue
    const unsigned crank = MC ? i8_cta_rank() : 0u;        // MC: the row half this CTA owns
    const int hfix = MC ? (int)crank : -1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long X = MC ? (blockIdx.x >> 1) : blockIdx.x;
    const int split = blockIdx.y;
    const int jt_begin = p.split_begin[split], jt_end = p.split_begin[split + 1];

    for (int i = threadIdx.x; i < jt_end - jt_begin; i += V2_THREADS) ttab[i] = v2_tile_pack(p, jt_begin + i);
    if (threadIdx.x == 0) {
        // MC: a slot is free once BOTH CTAs have consumed it (either one's multicast writes into both)
        for (int s = 0; s < STAGES; ++s) { i8_mbar_init(full0 + 8 * s, 1); i8_mbar_init(empty0 + 8 * s, MC ? 2 : 1); }
        for (int b = 0; b < 2; ++b) {
            i8_mbar_init(accfull0 + 8 * b, 1);
            i8_mbar_init(accempty0 + 8 * b, V2_EPI_WARPS);
        }
        i8_mbar_init(gobar, V2_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == V2_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (p.debug & 8) {                    // profiling: operands all zero (with debug & 4: the MMAs then run on zeros)
        for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += V2_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC) i8_cluster_sync();            // the peer's barriers must exist before a multicast copy or commit can signal them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    if (warp == V2_EPI_WARPS) {
        // ------------------------------------------------------------------ producer (whole warp walks the loop, one lane issues)
        unsigned slot = 0, ph = 0;
        v2_for_units<SOUTER>(ttab, T, jt_begin, jt_end, hfix, [&](const V2Tile& t, int h, int s) {
            // xi: [X tile][global k-block][8 KiB]; digits of block row (c, J): [trait s][row half h][kb][14 KiB]
            const signed char* asrc = p.xi + (X * p.nkb_total + (t.prow0 - 64 * t.J) / 64) * (long long)V2_A_BYTES;
            const signed char* gsrc = p.G + t.gblk * V2_BLK_BYTES + (long long)((s * 2 + h) * (t.J + 1)) * V2_B_BYTES;
            for (int kb = 0; kb < t.nkb; kb += V2_KS) {
                const unsigned nk = (unsigned)min(V2_KS, t.nkb - kb);
                if (p.debug & 4) continue;                   // profiling: no loads (the MMA warp then skips its `full` waits)
                i8_mbar_wait(empty0 + 8 * slot, ph ^ 1u);
                if (v2_elect()) {
                    const unsigned dst = i8_smem_u32(smem) + slot * STAGE_BYTES, bar = full0 + 8 * slot;
                    i8_mbar_expect_tx(bar, nk * (V2_A_BYTES + V2_B_BYTES));
                    if (MC) {       // k-block (kb + rank) of the xi tile goes to both CTAs of the pair
                        if (crank < nk) v2_bulk_g2s_mc(dst + crank * V2_A_BYTES, asrc + (long long)(kb + crank) * V2_A_BYTES, V2_A_BYTES, bar, 3);
                    } else {
                        i8_bulk_g2s(dst, asrc + (long long)kb * V2_A_BYTES, nk * V2_A_BYTES, bar);
                    }
                    i8_bulk_g2s(dst + V2_B_OFF, gsrc + (long long)kb * V2_B_BYTES, nk * V2_B_BYTES, bar);
                }
                __syncwarp();
                if (++slot == STAGES) { slot = 0; ph ^= 1u; }
            }
        });
    } else if (warp == V2_EPI_WARPS + 1) {
        // ---------------------------------------------------------------------- MMA issuer
        const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(V2_M >> 4) << 24) | ((unsigned)(V2_N >> 3) << 17);
        constexpr unsigned B_LBO = (V2_N / 8) * 128;                             // k-chunk stride of the digit operand
        // descriptors of slot 0, k-block 0, k32-step 0; everything else is an add on the 14-bit address field
        const unsigned long long da0 = i8_desc(i8_smem_u32(smem), V2_M * 16, 128);
        const unsigned long long db0 = i8_desc(i8_smem_u32(smem) + V2_B_OFF, B_LBO, 128);
        unsigned slot = 0, ph = 0, u = 0;
        v2_for_units<SOUTER>(ttab, T, jt_begin, jt_end, hfix, [&](const V2Tile& t, int, int) {
            const unsigned buf = u & 1u;
            i8_mbar_wait(accempty0 + 8 * buf, ((u >> 1) & 1u) ^ 1u);             // the epilogue has drained this accumulator
            if (GAP && u >= 1 && !(p.debug & 16)) i8_mbar_wait(gobar, (u - 1u) & 1u);   // no MMA in flight during an FP64 burst
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned tacc = tmem_base + buf * V2_BUF_COLS;
            for (int kb = 0; kb < t.nkb; kb += V2_KS) {
                const int nk = min(V2_KS, t.nkb - kb);
                if (!(p.debug & 4)) i8_mbar_wait(full0 + 8 * slot, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (v2_elect()) {
                    const unsigned long long das = da0 + ((slot * STAGE_BYTES) >> 4), dbs = db0 + ((slot * STAGE_BYTES) >> 4);
#pragma unroll
                    for (int q = 0; q < V2_KS; ++q) {
                        if (q < nk && !(p.debug & 2)) {
#pragma unroll
                            for (int kk = 0; kk < V2_KB / 32; ++kk) {
                                const unsigned long long da = das + ((q * V2_A_BYTES + kk * 2 * (V2_M * 16)) >> 4);
                                const unsigned long long db = dbs + ((q * V2_B_BYTES + kk * 2 * B_LBO) >> 4);
                                i8_mma(tacc, da, db, idesc, (kb | q | kk) != 0);
                            }
                        }
                    }
                    if (MC) v2_commit_mc(empty0 + 8 * slot, 3);                  // frees the slot when these MMAs retire
                    else i8_commit(empty0 + 8 * slot);
                    if (kb + V2_KS >= t.nkb) i8_commit(accfull0 + 8 * buf);      // accumulator of this unit complete
                }
                __syncwarp();
                if (++slot == STAGES) { slot = 0; ph ^= 1u; }
            }
            ++u;
        });
    } else {
        // ------------------------------------------------------------------ epilogue: thread = one cross x 8 SNP rows
        const int quarter = warp & 3, cg = warp >> 2;          // TMEM lane quarter is fixed by warp id % 4
        const int etid = threadIdx.x;                          // 0 .. 511
        const long long unit = X * V2_M + quarter * 32 + lane;
        const bool live = unit < p.n_units;
        int pa = 0, pb = 0;
        if (live) { pa = p.unit_a[unit]; pb = (p.mode == MASK_XI) ? p.unit_b[unit] : 0; }
        const unsigned tlane = tmem_base + ((unsigned)(quarter * 32) << 16) + cg * 8;
        // this thread's 8 xi values of the tile whose first row has padded index prow, as (nz, neg) bit fields. The loads are
        // issued a unit ahead (raw words in registers) and only combined when the unit is processed
        struct RawMask { unsigned a, b, a2, b2; };
        auto load_raw = [&](long long prow) -> RawMask {
            RawMask r{0u, 0u, 0u, 0u};
            if (live) {
                const long long wi = prow >> 5;
                r.a = __ldg(p.planeA + (long long)pa * p.wpp + wi); r.b = __ldg(p.planeB + (long long)pa * p.wpp + wi);
                if (p.mode == MASK_XI) { r.a2 = __ldg(p.planeA + (long long)pb * p.wpp + wi); r.b2 = __ldg(p.planeB + (long long)pb * p.wpp + wi); }
            }
            return r;
        };
        auto finish_masks = [&](const RawMask& r, unsigned& nz, unsigned& ng) {
            if (p.mode == MASK_XI) {
                nz = (r.a ^ r.b) & (r.a2 ^ r.b2);
                ng = ((~r.a & r.b) ^ (~r.a2 & r.b2)) & nz;
            } else {
                nz = r.a ^ r.b;
                ng = (p.mode == MASK_DELTA) ? (~r.a & r.b) : 0u;
            }
            nz = (nz >> (8 * cg)) & 0xFFu;
            ng = (ng >> (8 * cg)) & 0xFFu;
        };
        auto load_masks = [&](long long prow, unsigned& nz, unsigned& ng) { finish_masks(load_raw(prow), nz, ng); };
        // one accumulator: 7 digits x this thread's 8 rows, four rows at a time: acc(b, W[4]) gets the exact 64-bit sums of rows
        // 4b .. 4b+3; the accumulator is released as soon as its last column is in registers
        auto drain = [&](unsigned u, auto acc) {
            const unsigned buf = u & 1u;
            i8_mbar_wait(accfull0 + 8 * buf, (u >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned tl = tlane + buf * V2_BUF_COLS;
            if (p.debug & 1) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) i8_mbar_arrive(accempty0 + 8 * buf);
                return;
            }
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                long long W[4];
                {
                    int d[V2_DIG][4];
                    if (p.debug & 64) {                        // profiling: no TMEM loads
#pragma unroll
                        for (int i = 0; i < V2_DIG; ++i)
#pragma unroll
                            for (int c = 0; c < 4; ++c) d[i][c] = (int)u + i + c;
                    } else {
#pragma unroll
                        for (int i = 0; i < V2_DIG; ++i) v2_tmem_ld4(tl + i * V2_ROWS + b * 4, d[i]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                    if (b == 1) {                              // accumulator fully read: the MMAs of unit u + 2 may overwrite it
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) i8_mbar_arrive(accempty0 + 8 * buf);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) W[c] = v2_combine(d[0][c], d[1][c], d[2][c], d[3][c], d[4][c], d[5][c], d[6][c]);
                }
                if (!(p.debug & 32)) acc(b, W);                // profiling: TMEM loads only
                else if (W[0] + W[1] + W[2] + W[3] == 0x7fffffffffffffffLL) __trap();
            }
        };
        // the FP64 part runs only over the rows where this thread's xi is non-zero (13 % of them for a cross): nib = the 4 xi
        // bits of rows 4b .. 4b+3, f(e, W_e) per set bit (a divergent loop; its trip count is the warp's maximum popcount)
        auto for_nonzero = [&](unsigned nib, int b, const long long (&W)[4], auto f) {
            while (nib) {
                const int c = __ffs(nib) - 1;
                nib &= nib - 1u;
                const long long w = (c & 2) ? ((c & 1) ? W[3] : W[2]) : ((c & 1) ? W[1] : W[0]);
                f(b * 4 + c, w);
            }
        };
        if constexpr (!SOUTER) {
            double Z[TT][TT];
#pragma unroll
            for (int t = 0; t < TT; ++t)
#pragma unroll
                for (int s = 0; s < TT; ++s) Z[t][s] = 0.0;
            // row-data buffer: e_t[j] (T x 32 doubles), then per (s, j) the pair (m, cm) = (2^-49 scale_s[j], -2^52 m)
            constexpr int RB = 3 * TT * V2_ROWS;
            auto load_row = [&](long long prow) -> double {
                if (etid >= RB) return 0.0;
                const int which = etid / (TT * V2_ROWS), rem = etid - which * (TT * V2_ROWS), t = rem >> 5, j = rem & 31;
                const double v = __ldg((which ? p.scale : p.emat) + (long long)t * p.mp_total + prow + j);
                return which == 0 ? v : (which == 1 ? v * 0x1p-49 : v * -8.0);
            };
            auto row_slot = [&]() -> int {                     // where this thread's value goes inside a buffer
                const int which = etid / (TT * V2_ROWS), rem = etid - which * (TT * V2_ROWS);
                return which == 0 ? rem : TT * V2_ROWS + 2 * rem + (which - 1);
            };
            unsigned u = 0, ti = 0;
            int jt = jt_begin, h = MC ? hfix : 0;
            unsigned nz_n = 0u, ng_n = 0u;
            double row_n = 0.0;
            V2Tile cur{};
            if (jt < jt_end) {
                cur = v2_tile_unpack(ttab[jt - jt_begin]);
                row_n = load_row(cur.prow0 + V2_ROWS * h);
                load_masks(cur.prow0 + V2_ROWS * h, nz_n, ng_n);
                if (etid < RB) rowbuf[row_slot()] = row_n;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(V2_EPI_THREADS) : "memory");
            while (jt < jt_end) {
                const unsigned nz = nz_n, ng = ng_n;
                // prefetch the next tile's row data and haplotype words; they land while this tile's T units are processed
                int jt2 = jt, h2 = h + 1;
                V2Tile nxt = cur;
                if (MC || h2 >= cur.nh) { jt2 = jt + 1; h2 = MC ? hfix : 0; if (jt2 < jt_end) nxt = v2_tile_unpack(ttab[jt2 - jt_begin]); }
                if (jt2 < jt_end) {
                    const long long prow2 = nxt.prow0 + V2_ROWS * h2;
                    row_n = load_row(prow2);
                    load_masks(prow2, nz_n, ng_n);
                }
                const double* rb = rowbuf + (ti & 1u) * RB;
                const double* erow = rb + cg * 8;                                   // e_t[j]: erow[t * 32 + e]
                const double2* mrow = reinterpret_cast<const double2*>(rb + TT * V2_ROWS) + cg * 8;     // (m, cm): mrow[s * 32 + e]
#pragma unroll
                for (int s = 0; s < TT; ++s, ++u) {                                 // unrolled: Z[t][s] must stay in registers
                    drain(u, [&](int b, const long long (&W)[4]) {
                        for_nonzero((nz >> (4 * b)) & 0xFu, b, W, [&](int e, long long w) {
                            const unsigned sgn = (ng << (31 - e)) & 0x80000000u;
                            const double2 mc = mrow[s * V2_ROWS + e];
                            const double vm = v2_scaled(w, v2_flip(mc.x, sgn), v2_flip(mc.y, sgn));
#pragma unroll
                            for (int t = 0; t < TT; ++t) Z[t][s] = fma(erow[t * V2_ROWS + e], vm, Z[t][s]);
                        });
                    });
                }
                if (jt2 < jt_end && etid < RB) rowbuf[((ti + 1u) & 1u) * RB + row_slot()] = row_n;
                asm volatile("bar.sync 1, %0;" ::"n"(V2_EPI_THREADS) : "memory");
                jt = jt2; h = h2; cur = nxt; ++ti;
            }
            if (live) {
                double* out = p.Zpart + ((long long)((split * (MC ? 2 : 1) + (int)crank) * V2_NPART + cg) * p.n_units + unit) * (T * T);
#pragma unroll
                for (int t = 0; t < TT; ++t)
#pragma unroll
                    for (int s = 0; s < TT; ++s) out[t * T + s] = Z[t][s];
            }
        } else if constexpr (IA) {
            // ---- trait s outer, tile inner, exact integer accumulation. Per (t, s) the weights c[j] = e_t[j] scale_s[j] are
            // 63-bit fixed point with ONE exponent for all SNPs (i8_build_qtab_kernel): q = q2 2^42 + q1 2^21 + q0; a
            // non-zero row contributes xi W q with W = w2 2^42 + w1 2^21 + w0 the exact 60-bit sum above. The six partial
            // products of weight >= 2^42 go to three int64 accumulators per t (what is dropped is < 2^-60 of the largest
            // product); they cannot overflow below ~2^20 non-zero rows per cross and launch. FP64 appears once, at the end.
            constexpr int M21 = 0x1FFFFF;
            const int QB = V2_ROWS * T;                        // int4 per q-tile buffer: [row j][trait t] = (q0, q1, q2, -)
            int4* qbuf = reinterpret_cast<int4*>(rowbuf);
            auto q_src = [&](int s, long long prow) { return p.qtab + ((long long)s * p.mp_total + prow) * T + etid; };
            // walk (s, tile) one step ahead of the unit being processed
            int s_n = 0, jt_n = jt_begin, h_n = MC ? hfix : 0;
            V2Tile t_n{};
            bool valid_n = jt_begin < jt_end && T > 0;
            if (valid_n) t_n = v2_tile_unpack(ttab[jt_n - jt_begin]);
            auto advance = [&]() {
                if (!MC && h_n + 1 < t_n.nh) { ++h_n; return; }
                h_n = MC ? hfix : 0;
                if (++jt_n >= jt_end) { jt_n = jt_begin; if (++s_n >= T) { valid_n = false; return; } }
                t_n = v2_tile_unpack(ttab[jt_n - jt_begin]);
            };
            int4 q_n = make_int4(0, 0, 0, 0);
            RawMask raw_n{0u, 0u, 0u, 0u};
            if (valid_n) {
                const long long prow = t_n.prow0 + V2_ROWS * h_n;
                if (etid < QB) qbuf[etid] = __ldg(q_src(s_n, prow));
                raw_n = load_raw(prow);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(V2_EPI_THREADS) : "memory");
            constexpr int V2_SLOTS = 4;
            unsigned u = 0;
            for (int s = 0; s < T; ++s) {
                long long a2[TT], a3[TT], a4[TT];
#pragma unroll
                for (int t = 0; t < TT; ++t) { a2[t] = 0; a3[t] = 0; a4[t] = 0; }
                // one non-zero row: W with the sign of xi, as balanced base-2^21 words, times the T weights of its row
                auto accumulate = [&](long long w, const int4* qr) {
                    w += (1ll << 20) + (1ll << 41);
                    const int w0 = ((int)w & M21) - (1 << 20), w1 = ((int)(w >> 21) & M21) - (1 << 20), w2 = (int)(w >> 42);
#pragma unroll
                    for (int t = 0; t < TT; ++t) {
                        if (TT == T || t < T) {
                            const int4 q = qr[t];
                            a4[t] = v2_madw(w2, q.z, a4[t]);
                            a3[t] = v2_madw(w2, q.y, v2_madw(w1, q.z, a3[t]));
                            a2[t] = v2_madw(w2, q.x, v2_madw(w1, q.y, v2_madw(w0, q.z, a2[t])));
                        }
                    }
                };
                while (valid_n && s_n == s) {
                    unsigned nz, ng;
                    finish_masks(raw_n, nz, ng);
                    advance();                                  // prefetch the next unit's weights and haplotype words
                    if (valid_n) {
                        const long long prow = t_n.prow0 + V2_ROWS * h_n;
                        if (etid < QB) q_n = __ldg(q_src(s_n, prow));
                        raw_n = load_raw(prow);
                    }
                    const int4* qb = qbuf + (u & 1u) * QB + cg * 8 * T;
                    long long W[8];
                    drain(u, [&](int b, const long long (&Wb)[4]) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) W[4 * b + c] = Wb[c];
                    });
                    auto pick = [&](int c) -> long long {
                        const long long a = (c & 1) ? W[1] : W[0], b2 = (c & 1) ? W[3] : W[2], c2 = (c & 1) ? W[5] : W[4], d2 = (c & 1) ? W[7] : W[6];
                        const long long lo = (c & 2) ? b2 : a, hi = (c & 2) ? d2 : c2;
                        const long long w = (c & 4) ? hi : lo;
                        return ((ng >> c) & 1u) ? -w : w;
                    };
                    // the first V2_SLOTS non-zero rows without divergence (an empty slot adds 0 x the weights of row 0) ...
                    unsigned m = nz;
#pragma unroll
                    for (int k = 0; k < V2_SLOTS; ++k) {
                        const bool ok = m != 0u;
                        const int c = ok ? __ffs(m) - 1 : 0;
                        m &= m - 1u;
                        accumulate(ok ? pick(c) : 0ll, qb + c * T);
                    }
                    while (m) {                                 // ... a fifth, sixth .. among 8: rare for a cross (0.3 %)
                        const int c = __ffs(m) - 1;
                        m &= m - 1u;
                        accumulate(pick(c), qb + c * T);
                    }
                    if (valid_n && etid < QB) qbuf[((u + 1u) & 1u) * QB + etid] = q_n;
                    asm volatile("bar.sync 1, %0;" ::"n"(V2_EPI_THREADS) : "memory");
                    ++u;
                }
                if (live) {
                    double* out = p.Zpart + ((long long)((split * (MC ? 2 : 1) + (int)crank) * V2_NPART + cg) * p.n_units + unit) * (T * T);
#pragma unroll
                    for (int t = 0; t < TT; ++t)
                        if (t < T) {
                            const double v = fma((double)a4[t], 0x1p21, (double)a3[t]);
                            out[t * T + s] = ldexp(fma(v, 0x1p21, (double)a2[t]), p.gexp[t * T + s] - 62 - 49 + 42);
                        }
                }
            }
        } else {
            // ---- trait s outer, tile inner: Z[:, s] in registers, any T <= TT. FP64 instructions stall the tensor pipe (about
            // 4.5 SM cycles each when they trickle in, 0.85 back to back: tools/tc_contention_probe.cu), so a unit is handled
            // in two phases: an integer phase (TMEM loads, exact 64-bit sums, the first V2_SLOTS non-zero rows of this thread
            // moved to fixed registers - no divergence) and one short FP64 burst that all 16 warps enter together.
            constexpr int V2_SLOTS = 4;
            // row data of a unit, PER WARP (its 8 rows): per row (m, cm, e_0 .. e_{T-1}) with (m, cm) = (2^-49 scale_s[j], -2^52 m),
            // RW doubles per row; two buffers per warp, touched by that warp only - the epilogue warps never meet at a CTA-wide
            // barrier. scale_s[j] is a power of two, so m and cm are made with integer instructions.
            constexpr int RW = (TT + 2 + 1) & ~1;               // doubles per row (even: 16-byte loads)
            constexpr int RBW = 8 * RW;
            constexpr int NV = ((TT + 1) * 8 + 31) / 32;        // loads per lane: column c = i >> 3 (c < T: e_c, c == T: scale_s), row i & 7
            const unsigned wbuf = i8_smem_u32(rowbuf) + warp * (2 * RBW * 8);
            auto row_src = [&](int s, long long prow, int i) -> double {
                const int c = i >> 3, r = i & 7;
                return __ldg((c < T ? p.emat + (long long)c * p.mp_total : p.scale + (long long)s * p.mp_total) + prow + cg * 8 + r);
            };
            auto row_store = [&](unsigned buf, int i, double v) {
                const int c = i >> 3, r = i & 7;
                const unsigned a = buf + (r * RW) * 8;
                if (c < T) v2_sts64(a + (2 + c) * 8, v);
                else if (c == T) {
                    const int hi = __double2hiint(v);
                    const bool ok = ((hi >> 20) & 0x7FF) > 64;                    // anything smaller counts as a zero row
                    v2_sts64(a, ok ? __hiloint2double(hi - (49 << 20), 0) : 0.0);
                    v2_sts64(a + 8, ok ? __hiloint2double((hi + (3 << 20)) | (int)0x80000000, 0) : 0.0);
                }
            };
            int s_n = 0, jt_n = jt_begin, h_n = MC ? hfix : 0;
            V2Tile t_n{};
            bool valid_n = jt_begin < jt_end && T > 0;
            if (valid_n) t_n = v2_tile_unpack(ttab[jt_n - jt_begin]);
            auto advance = [&]() {
                if (!MC && h_n + 1 < t_n.nh) { ++h_n; return; }
                h_n = MC ? hfix : 0;
                if (++jt_n >= jt_end) { jt_n = jt_begin; if (++s_n >= T) { valid_n = false; return; } }
                t_n = v2_tile_unpack(ttab[jt_n - jt_begin]);
            };
            double row_n[NV];
            RawMask raw_n{0u, 0u, 0u, 0u};
            if (valid_n) {
                const long long prow = t_n.prow0 + V2_ROWS * h_n;
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (lane + 32 * v < (T + 1) * 8) row_store(wbuf, lane + 32 * v, row_src(s_n, prow, lane + 32 * v));
                raw_n = load_raw(prow);
            }
            __syncwarp();
            // the non-zero rows of the PREVIOUS unit wait in registers for their FP64 burst
            long long sw[V2_SLOTS];
            unsigned sa[V2_SLOTS], ssg[V2_SLOTS];               // shared address of the row's data, sign of xi
#pragma unroll
            for (int k = 0; k < V2_SLOTS; ++k) { sw[k] = 0ll; sa[k] = wbuf; ssg[k] = 0u; }
            unsigned u = 0;
            for (int s = 0; s < T; ++s) {
                double zs[TT];
#pragma unroll
                for (int t = 0; t < TT; ++t) zs[t] = 0.0;
                auto item = [&](long long w, unsigned a, unsigned sg) {
                    const double2 mc = v2_lds128(a);
                    const double vm = v2_scaled(w, v2_flip(mc.x, sg), v2_flip(mc.y, sg));
#pragma unroll
                    for (int t2 = 0; t2 < TT; t2 += 2) {
                        const double2 e = v2_lds128(a + (2 + t2) * 8);
                        if (t2 < T) zs[t2] = fma(e.x, vm, zs[t2]);
                        if (t2 + 1 < TT && t2 + 1 < T) zs[t2 + 1] = fma(e.y, vm, zs[t2 + 1]);
                    }
                };
                auto burst = [&]() {
#pragma unroll
                    for (int k = 0; k < V2_SLOTS; ++k) item(sw[k], sa[k], ssg[k]);
                };
                while (valid_n && s_n == s) {
                    unsigned nz, ng;
                    finish_masks(raw_n, nz, ng);
                    advance();                                  // prefetch the next unit's row data and haplotype words
                    if (valid_n) {
                        const long long prow = t_n.prow0 + V2_ROWS * h_n;
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            if (lane + 32 * v < (T + 1) * 8) row_n[v] = row_src(s_n, prow, lane + 32 * v);
                        raw_n = load_raw(prow);
                    }
                    const unsigned rb = wbuf + (u & 1u) * (RBW * 8);
                    const unsigned buf = u & 1u;
                    i8_mbar_wait(accfull0 + 8 * buf, (u >> 1) & 1u);     // MMA(u) is complete and MMA(u + 1) has not been issued:
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    burst();                                             // the tensor pipe is idle while the FP64 burst of unit u - 1 runs
                    __syncwarp();
                    if (lane == 0) i8_mbar_arrive(gobar);                // MMA(u + 1) may start; the integer phase below overlaps it
                    long long W[8];
                    const unsigned tl = tlane + buf * V2_BUF_COLS;
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        int d[V2_DIG][4];
#pragma unroll
                        for (int i = 0; i < V2_DIG; ++i) v2_tmem_ld4(tl + i * V2_ROWS + b * 4, d[i]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (b == 1) {
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) i8_mbar_arrive(accempty0 + 8 * buf);
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) W[4 * b + c] = v2_combine(d[0][c], d[1][c], d[2][c], d[3][c], d[4][c], d[5][c], d[6][c]);
                    }
                    auto pick = [&](int c) -> long long {
                        const long long a = (c & 1) ? W[1] : W[0], b2 = (c & 1) ? W[3] : W[2], c2 = (c & 1) ? W[5] : W[4], d2 = (c & 1) ? W[7] : W[6];
                        const long long lo = (c & 2) ? b2 : a, hi = (c & 2) ? d2 : c2;
                        return (c & 4) ? hi : lo;
                    };
                    unsigned m = nz;
#pragma unroll
                    for (int k = 0; k < V2_SLOTS; ++k) {               // no divergence: an empty slot carries W = 0
                        const bool ok = m != 0u;
                        const int c = ok ? __ffs(m) - 1 : 0;
                        m &= m - 1u;
                        sw[k] = ok ? pick(c) : 0ll;
                        sa[k] = rb + c * (RW * 8);
                        ssg[k] = (ng << (31 - c)) & 0x80000000u;
                    }
                    while (m) {                                 // a fifth, sixth .. non-zero row among 8: rare for a cross (0.3 %)
                        const int c = __ffs(m) - 1;
                        m &= m - 1u;
                        item(pick(c), rb + c * (RW * 8), (ng << (31 - c)) & 0x80000000u);
                    }
                    // unit u + 1's data go where unit u - 1's were: this warp has finished the burst of u - 1 above
                    if (valid_n) {
                        const unsigned nb = wbuf + ((u + 1u) & 1u) * (RBW * 8);
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            if (lane + 32 * v < (T + 1) * 8) row_store(nb, lane + 32 * v, row_n[v]);
                    }
                    __syncwarp();
                    ++u;
                }
                burst();                                        // the last unit of this trait (the next trait's MMAs are already running)
#pragma unroll
                for (int k = 0; k < V2_SLOTS; ++k) sw[k] = 0ll;
                if (live) {
                    double* out = p.Zpart + ((long long)((split * (MC ? 2 : 1) + (int)crank) * V2_NPART + cg) * p.n_units + unit) * (T * T);
#pragma unroll
                    for (int t = 0; t < TT; ++t)
                        if (t < T) out[t * T + s] = zs[t];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC) i8_cluster_sync();            // no CTA may exit while its peer can still multicast into it
    if (warp == V2_EPI_WARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ prep: fixed-point row weights
// gexp[t * T + s] = E with max_j |e_t[j] scale_s[j]| < 2^E (0 if all are zero)
__global__ void i8_qexp_kernel(const double* __restrict__ emat, const double* __restrict__ scale, long long mp_total, int T,
                               int* __restrict__ gexp) {
    const int t = blockIdx.x / T, s = blockIdx.x - t * T;
    double mx = 0.0;
    for (long long j = threadIdx.x; j < mp_total; j += blockDim.x)
        mx = fmax(mx, fabs(emat[(long long)t * mp_total + j] * scale[(long long)s * mp_total + j]));
    __shared__ double sh[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, sh[i]);
        int e = 0;
        if (mx > 0.0) (void)frexp(mx, &e);
        gexp[blockIdx.x] = e;
    }
}
// qtab[(s * mp + j) * T + t] = (q0, q1, q2, 0): q = rint(e_t[j] scale_s[j] 2^(62 - E)) = q2 2^42 + q1 2^21 + q0, q0, q1 in [-2^20, 2^20)
__global__ void i8_qtab_kernel(const double* __restrict__ emat, const double* __restrict__ scale, long long mp_total, int T,
                               const int* __restrict__ gexp, int4* __restrict__ qtab) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;          // (s, j, t), t fastest
    if (idx >= (long long)T * mp_total * T) return;
    const int t = (int)(idx % T);
    const long long sj = idx / T, j = sj % mp_total;
    const int s = (int)(sj / mp_total);
    const double c = emat[(long long)t * mp_total + j] * scale[(long long)s * mp_total + j];
    const long long q = __double2ll_rn(ldexp(c, 62 - gexp[t * T + s])) + (1ll << 20) + (1ll << 41);     // balanced words
    qtab[idx] = make_int4(((int)q & 0x1FFFFF) - (1 << 20), ((int)(q >> 21) & 0x1FFFFF) - (1 << 20), (int)(q >> 42), 0);
}
cudaError_t launch_i8_build_qtab(const double* emat, const double* scale, long long mp_total, int T, int* gexp, int4* qtab,
                                 cudaStream_t st) {
    i8_qexp_kernel<<<T * T, 1024, 0, st>>>(emat, scale, mp_total, T, gexp);
    const long long total = (long long)T * mp_total * T;
    i8_qtab_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(emat, scale, mp_total, T, gexp, qtab);
    return cudaGetLastError();
}

int i8v2_max_tiles_per_split() { return V2_TTAB; }
int i8v2_npart(bool mc) { return mc ? 2 * V2_NPART : V2_NPART; }

size_t i8v2_smem_bytes(int T, bool ia) {
    // row data: integer variant 2 buffers of 32 x T int4; tile-outer FP64 2 buffers of 3 T x 32 doubles (T <= 5); trait-outer
    // FP64 2 buffers of (T + 2) x 8 doubles per epilogue warp
    size_t rowdata = ia && T <= 8 ? (size_t)2 * V2_ROWS * T * 16 : (size_t)2 * (16 + 2) * 8 * 8 * V2_EPI_WARPS;     // sized for the largest TT
    if (!ia && T <= 5) rowdata = rowdata > (size_t)2 * 3 * T * V2_ROWS * 8 ? rowdata : (size_t)2 * 3 * T * V2_ROWS * 8;
    return (size_t)V2_STAGES * V2_STAGE_BYTES + (2 * V2_MAX_STAGES + 6) * 8 + (size_t)V2_TTAB * 16 + rowdata;
}

// grid = (2 x ntiles, nsplit) in clusters of two CTAs: the pair shares one 128-cross tile, CTA rank = row half
cudaError_t launch_i8v2_contract(const I8Params& p, int ntiles, int nsplit, bool mc, bool ia, cudaStream_t st) {
    if (ntiles == 0) return cudaSuccess;
    if (p.T > 8) ia = false;
    const size_t smem = i8v2_smem_bytes(p.T, ia);
    cudaError_t e = cudaSuccess;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(mc ? 2 * ntiles : ntiles, nsplit, 1);
    cfg.blockDim = dim3(V2_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = mc ? 1 : 0;
#define GMS_V2_ONE(TV, SO, PR, IA)                                                                                      \
    {                                                                                                                   \
        e = cudaFuncSetAttribute(i8v2_kernel<TV, SO, PR, IA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e != cudaSuccess) return e;                                                                                 \
        e = cudaLaunchKernelEx(&cfg, i8v2_kernel<TV, SO, PR, IA>, p);                                                   \
        if (e != cudaSuccess) return e;                                                                                 \
    }
#define GMS_V2(TV, SO, IA) { if (mc) GMS_V2_ONE(TV, SO, true, IA) else GMS_V2_ONE(TV, SO, false, IA) }
    if (ia) {
        switch (p.T) {
            case 1: GMS_V2(1, true, true) break;
            case 2: GMS_V2(2, true, true) break;
            case 3: GMS_V2(3, true, true) break;
            case 4: GMS_V2(4, true, true) break;
            case 5: GMS_V2(5, true, true) break;
            case 6: GMS_V2(6, true, true) break;
            default: GMS_V2(8, true, true) break;
        }
    } else if (p.debug & 128) {         // the tile-outer variant (whole T x T table in registers), kept for comparison
        switch (p.T) {
            case 1: GMS_V2(1, false, false) break;
            case 2: GMS_V2(2, false, false) break;
            case 3: GMS_V2(3, false, false) break;
            case 4: GMS_V2(4, false, false) break;
            case 5: GMS_V2(5, false, false) break;
            default: return cudaErrorInvalidValue;
        }
    } else {
        if (p.T <= 2) GMS_V2(2, true, false)
        else if (p.T <= 4) GMS_V2(4, true, false)
        else if (p.T <= 5) GMS_V2(5, true, false)
        else if (p.T <= 8) GMS_V2(8, true, false)
        else if (p.T <= 12) GMS_V2(12, true, false)
        else if (p.T <= 16) GMS_V2(16, true, false)
        else return cudaErrorInvalidValue;
    }
#undef GMS_V2
#undef GMS_V2_ONE
    return cudaGetLastError();
}

}  // namespace gms
