// Int8 tensor-core evaluation of the quadratic-form tables (per parent and per cross): GMS_MODE_INT8_TENSOR, what
// GMS_MODE_AUTO resolves to.
//
// FP64 has no tcgen05 kind, but this contraction has a special shape: one operand is TERNARY.
//     C_s[j,x] = sum_k G_s[j,k] xi_x[k],   G_s[j,k] = L'[j,k] d_s[k]  (FP64),   xi in {-1,0,+1}
// xi is exact in int8. G_s is rounded row-wise to I8_FRAC = 49 fractional bits of a shared power-of-two row scale 2^E(j,s)
// with max_k |G_s[j,k]| <= 2^E and split into I8_DIG = 7 BALANCED radix-256 digits g_i in [-128, 127] (i8digits.cuh; the
// build variant -DGMS_I8_DIG=6 keeps 47 bits in six planes: 14 % fewer MMAs for 4 x the rounding error, tools only):
// G = 2^(E - I8_FRAC) sum_i g_i 256^(I8_DIG-1-i) + r,  |r| <= 2^(E-50) (round to nearest). Each digit plane times xi
// is an int8 x int8 -> int32 product whose sum over the k of a row is EXACT in int32 for any chromosome length this library
// accepts (|sum| <= 128 n for n non-zero xi entries), so the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators in TMEM) do the O(m^2) work at int8 rate. The epilogue recombines the digit sums exactly in 64-bit
// integer arithmetic (exact while n < 2^15; verify.cu sends units with n >= 2^14 on one chromosome to the FP64 path),
// builds the FP64 bit pattern of xi_x[j] W 2^(E-I8_FRAC) with integer instructions (i8digits.cuh: i8_cvt_bits - FP64
// instructions stall the MMAs of the SM, integer ones do not), multiplies by d_t[j] and accumulates the T x T quadratic
// forms per cross in registers (TMEM lane = cross, so the reduction over SNPs j is thread-local: no shuffles, no atomics).
// The only FP64 instructions left are those T FMAs per (cross, row, trait).
// WHICH BLOCKS A ROW TILE SUMS: the matrix is symmetric and the table is X = Z + Z', so each unordered pair of 64-blocks can
// be evaluated in either orientation; common.cuh deals them evenly (row tile J owns blocks J, J + 1, ... modulo the
// chromosome's block count, about half of them) instead of the lower block triangle's 1 .. nb blocks per row tile: every
// accumulator then carries the same MMA work, longer than the epilogue of the previous one.
// WHAT THE EPILOGUE COSTS (measured with the -DGMS_PROFILE switches, 50k cassava crosses, 7 planes): MMAs + loads alone
// 27.3 ms; + the TMEM drain 31.5 (the one accumulator fills TMEM, so MMAs and drain alternate); + the integer conversion
// 32.6; + the FP64 FMAs 37.0 when they run in the window where the tensor pipe is idle anyway (GMS_I8_LATE, below), 39.1
// when they run next to the next unit's MMAs. Tried and dropped (all parity-green): compacting each thread's non-zero rows
// through shared memory so that a warp loops to its largest count (a third of the FP64 instructions, but 45.8 ms: the
// gathers spread the burst out); releasing the accumulator columns of the first MMA before the rest is read (+-0);
// skipping rows no cross of a warp needs (predication breaks the dense burst: slower).
// THE RESULT IS FIXED POINT RELATIVE TO THE ROW SCALE, not floating point: verify.cu bounds the error of every unit
// a posteriori and api.cu re-evaluates the units that fail the bound with the FP64 contraction (contract.cu).
//
// Orientation: D[M = 128 crosses (TMEM lanes)][N = 64 SNP rows j] += Xi[128 x K] * Gdigit_i[64 x K]^T,
// both operands K-major, no swizzle, pre-tiled in HBM in the UMMA canonical "interleaved" layout
// (8 rows x 16 bytes core matrices) so that one cp.async.bulk lands them ready for the descriptors.
// Warp roles: 0-7 epilogue (tcgen05.ld; two warps per TMEM lane quarter, 32 SNP rows each), 8 bulk-copy producer,
// 9 MMA issuer; by default two CTAs form one M = 256 tile (tcgen05 cta_group::2). DESIGN.md section 4 has the measurements
// behind every choice.
#include "kernels.cuh"
#include "i8ptx.cuh"

// profiling switches (results are garbage): only in a -DGMS_PROFILE build, never in the shipped library
#ifdef GMS_PROFILE
#define I8_DBG(bit) (p.debug & (bit))
#else
#define I8_DBG(bit) false
#endif

namespace gms {

constexpr int I8_M = 128;        // crosses per tile (TMEM lanes)
constexpr int I8_N = 64;         // SNP rows j per accumulator
constexpr int I8_KB = 64;        // k per pipeline stage (bytes of int8)
// I8_DIG = 7 (or 6) balanced radix-256 digits (i8digits.cuh)
constexpr int I8_NA = 4 * I8_N;                     // N of the first MMA of a k32-step: digits 0-3 (256 = the UMMA maximum)
constexpr int I8_NB = (I8_DIG - 4) * I8_N;          // N of the second: the remaining digits (192 or 128)
constexpr int I8_A_BYTES = I8_M * I8_KB;            // 8 KiB
constexpr int I8_B_BYTES = I8_N * I8_KB;            // 4 KiB per digit
constexpr int I8_B_ALL = I8_DIG * I8_B_BYTES;       // 28 (24) KiB: the 448- (384-) row digit operand of one k-block
static_assert(I8_NB % 32 == 0 && I8_NA + I8_NB <= 512, "accumulator layout");
// per-variant stage geometry: single-CTA UMMA keeps the whole digit operand, a cta_group::2 pair keeps half per CTA
template <bool CG2> struct I8Geo {
    static constexpr int B_STAGE = CG2 ? I8_B_ALL / 2 : I8_B_ALL;
    static constexpr int STAGE_BYTES = I8_A_BYTES + B_STAGE;           // 36 / 22 KiB (7 planes), 32 / 20 KiB (6)
};
constexpr int I8_MAX_STAGES = 10;
constexpr int I8_TTAB = 512;        // 64-row tiles per split that fit the shared-memory copy of the work list (the host plan keeps splits below it)
constexpr int I8_EPI_WARPS = 8;     // two per TMEM lane quarter, each takes 32 of the 64 SNP rows of a tile (16 warps: 96 registers
                                    // per thread and x4 TMEM loads, measured 13 % slower at T = 5)
constexpr int I8_RPT = 32;          // SNP rows per epilogue thread and tile = one word of the bit-planes
constexpr int I8_THREADS = 32 * (I8_EPI_WARPS + 2);

// ------------------------------------------------------------------------------------------ prep: xi images
// image (X tile, global k-block) : [chunk cc (4)][r1 (16)][r0 (8)][16 bytes], row r = 8 r1 + r0, k = 16 cc + b
__global__ void i8_build_xi_kernel(const unsigned* __restrict__ planeA, const unsigned* __restrict__ planeB,
                                   long long wpp, const int* __restrict__ sire, const int* __restrict__ dam,
                                   long long n_units, int mode, long long nkb_total, signed char* __restrict__ xi) {
    // one thread = one (unit, 64-k block): the unit's 64 mask entries come from two words of each bit-plane (one 8-byte load per
    // plane and parent) and fill the block's four 16-byte chunks. (One thread per 16-byte chunk re-read the same words four
    // times through 32-byte sectors: 0.95 ms per 50k cassava crosses for 1.7 GB written.)
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npad = (n_units + 2 * I8_M - 1) / (2 * I8_M) * (2 * I8_M);     // an even number of 128-cross tiles (CTA pairs)
    if (gid >= npad * nkb_total) return;
    // consecutive threads -> consecutive rows of one (X, k-block) tile, so that the 16-byte stores of a warp are contiguous
    const long long tile = gid / I8_M;                  // (X, kbglobal)
    const int r = (int)(gid - tile * I8_M);
    const long long X = tile / nkb_total;
    const long long kbg = tile - X * nkb_total;
    const long long unit = X * I8_M + r;
    unsigned nzw[2] = {0u, 0u}, ngw[2] = {0u, 0u};
    if (unit < n_units) {
        // words 2 kbg, 2 kbg + 1 of the parent's planes: 8-byte aligned (wpp is a multiple of 4)
        const long long w = kbg * 2;
        const int pa = sire[unit];
        const uint2 A = *reinterpret_cast<const uint2*>(planeA + (long long)pa * wpp + w);
        const uint2 B = *reinterpret_cast<const uint2*>(planeB + (long long)pa * wpp + w);
        if (mode == MASK_XI) {
            const int pb = dam[unit];
            const uint2 A2 = *reinterpret_cast<const uint2*>(planeA + (long long)pb * wpp + w);
            const uint2 B2 = *reinterpret_cast<const uint2*>(planeB + (long long)pb * wpp + w);
            nzw[0] = (A.x ^ B.x) & (A2.x ^ B2.x);
            nzw[1] = (A.y ^ B.y) & (A2.y ^ B2.y);
            ngw[0] = ((~A.x & B.x) ^ (~A2.x & B2.x)) & nzw[0];
            ngw[1] = ((~A.y & B.y) ^ (~A2.y & B2.y)) & nzw[1];
        } else {
            nzw[0] = A.x ^ B.x;
            nzw[1] = A.y ^ B.y;
            ngw[0] = (mode == MASK_DELTA) ? (~A.x & B.x) : 0u;
            ngw[1] = (mode == MASK_DELTA) ? (~A.y & B.y) : 0u;
        }
    }
    const int r1 = r >> 3, r0 = r & 7;
    signed char* tile_base = xi + (X * nkb_total + kbg) * (long long)I8_A_BYTES;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {                    // chunk cc = k 16 cc .. 16 cc + 15 of the block
        const unsigned nz = nzw[cc >> 1] >> (16 * (cc & 1)), ng = ngw[cc >> 1] >> (16 * (cc & 1));
        unsigned out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int bit = 4 * q + b;
                const unsigned byte = ((nz >> bit) & 1u) ? (((ng >> bit) & 1u) ? 0xFFu : 0x01u) : 0u;
                v |= byte << (8 * b);
            }
            out[q] = v;
        }
        *reinterpret_cast<uint4*>(tile_base + ((cc * 16 + r1) * 8 + r0) * 16) = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

// ------------------------------------------------------------------------------------------ prep: G digits
// element (row, col) of the chromosome's L' (lower block triangle at 64 granularity, diagonal 64 x 64
// blocks halved) read from the FP64 tile image, whose own convention is 128-granular (common.cuh)
__device__ __forceinline__ double lprime64(const double* __restrict__ tiles, const ChromDesc& cd, int row, int col) {
    const int I = row >> 7, kc = col >> 5;
    const long long ch = chunks_before_rowtile(I) + kc;
    const int mb = (row & 127) >> 4, hi = (row & 15) >> 3, lane = ((row & 7) << 2) | (col & 3), q = (col & 31) >> 2;
    double v = tiles[cd.tile_off + ch * CHUNK + (((q * 8 + mb) * 32 + lane) << 1) + hi];
    if ((col >> 7) == I) v *= 2.0;                     // undo the 128-granular halving
    if ((col >> 6) == (row >> 6)) v *= 0.5;            // apply the 64-granular one
    return v;
}
// element (row, col) of the chromosome's symmetric matrix for ANY pair of 64-blocks (diagonal blocks halved): blocks above
// the diagonal are read through their mirror image in the stored lower block triangle
__device__ __forceinline__ double lsym64(const double* __restrict__ tiles, const ChromDesc& cd, int row, int col) {
    return (col >> 6) > (row >> 6) ? lprime64(tiles, cd, col, row) : lprime64(tiles, cd, row, col);
}

// Row scales, two steps. (1) rowmax[s][pj] = max_k |G_s[j,k]| over row j's lower-triangle range: one CTA per 64 x 64 block
// with the digit builder's access pattern (every L' element read once, coalesced, used for all T traits), reduced in shared
// memory and merged with an atomic max on the bit pattern (non-negative doubles order like unsigned integers; the buffer
// starts at zero). (2) scale = 2^E with rowmax <= 0.996 * 2^E (1 for a zero row; a row below 2^-900 counts as zero), in place.
__global__ void i8_rowmax_kernel(const double* __restrict__ tiles, const ChromDesc* __restrict__ chrom, int c_begin, int c_end,
                                 const int* __restrict__ blkpref, const double* __restrict__ emat, long long mp_total, int T,
                                 unsigned long long* __restrict__ rowmax) {
    int c = c_begin;
    while (c + 1 < c_end && (int)blockIdx.x >= blkpref[c + 1]) ++c;
    const ChromDesc cd = chrom[c];
    const int blk = blockIdx.x - blkpref[c];
    const int nb = (cd.m + 63) >> 6;
    int J, bi;
    i8_block_decode(nb, blk, J, bi);                  // the bi-th block of row tile J (common.cuh) ...
    const int kb = i8_block_k(nb, J, bi);             // ... is k-block kb of the chromosome
    const int r = threadIdx.x & 63, cc = threadIdx.x >> 6;      // 256 threads: 64 rows x 4 chunks of 16 columns
    const int row = J * 64 + r;
    __shared__ unsigned long long sh[64][MAX_T];
    for (int i = threadIdx.x; i < 64 * MAX_T; i += blockDim.x) (&sh[0][0])[i] = 0ull;
    __syncthreads();
    if (row < cd.m) {
        double lp[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int col = kb * 64 + cc * 16 + b;
            lp[b] = col < cd.m ? fabs(lsym64(tiles, cd, row, col)) : 0.0;
        }
        const double* ecol = emat + cd.pad_off + kb * 64 + cc * 16;
        for (int s = 0; s < T; ++s) {
            double mx = 0.0;
#pragma unroll
            for (int b = 0; b < 16; ++b) mx = fmax(mx, lp[b] * fabs(ecol[(long long)s * mp_total + b]));
            atomicMax(&sh[r][s], (unsigned long long)__double_as_longlong(mx));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * T; i += blockDim.x) {
        const int rr = i % 64, s = i / 64;
        if (J * 64 + rr < cd.m && sh[rr][s]) atomicMax(rowmax + (long long)s * mp_total + cd.pad_off + J * 64 + rr, sh[rr][s]);
    }
}
__global__ void i8_scale_kernel(double* __restrict__ scale, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = scale[i];
    // rowmax <= I8_FILL * 2^e, so that rn(G 2^(I8_FRAC - e)) fits I8_DIG balanced radix-256 digits: no entry ever saturates
    scale[i] = ldexp(1.0, v > 0x1p-900 ? i8_scale_exp(v) : 0);
}

// digit images: (c, J, kb, s): one (I8_DIG x 64)-row K-major operand per (block, trait); one thread = 16 consecutive k of a
// row, L' read once and reused for all T traits
__global__ void i8_build_digits_kernel(const double* __restrict__ tiles, const ChromDesc* __restrict__ chrom, int c_begin,
                                       int c_end, const int* __restrict__ blkpref, const double* __restrict__ emat,
                                       long long mp_total, int T, const double* __restrict__ scale,
                                       const long long* __restrict__ gbase_c, signed char* __restrict__ G, int layout) {
    int c = c_begin;
    while (c + 1 < c_end && (int)blockIdx.x >= blkpref[c + 1]) ++c;     // blkpref[c] = first block of chromosome c
    const ChromDesc cd = chrom[c];
    const long long gbase = gbase_c[c];
    const int blk = blockIdx.x - blkpref[c];          // index of the block in the chromosome's list: row tile J, its bi-th block
    const int nb = (cd.m + 63) >> 6;
    int J, bi;
    i8_block_decode(nb, blk, J, bi);
    const int kb = i8_block_k(nb, J, bi);
    const int r = threadIdx.x & 63, cc = threadIdx.x >> 6;      // 256 threads: 64 rows x 4 chunks
    const int row = J * 64 + r;
    double lp[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const int col = kb * 64 + cc * 16 + b;
        lp[b] = (row < cd.m && col < cd.m) ? lsym64(tiles, cd, row, col) : 0.0;
    }
    // the I8_DIG digit planes of one (block, trait) form ONE K-major operand of I8_DIG x 64 rows:
    // [chunk cc (4)][row group R1 = 8 i + r1 (8 I8_DIG)][r0 (8)][16 bytes]  ->  LBO = 8 I8_DIG * 128 B, SBO = 128 B,
    // so a k32-step needs two MMAs (N = 256: digits 0-3, N = 192 or 128: the remaining ones) instead of one N = 64 MMA per plane
    const int r1 = r >> 3, r0 = r & 7;
    const double* ecol = emat + cd.pad_off + kb * 64 + cc * 16;
    for (int s = 0; s < T; ++s) {
        const double inv = 1.0 / scale[(long long)s * mp_total + cd.pad_off + row];     // exact: a power of two
        signed char d[I8_DIG][16];
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            // |x| <= I8_FILL (i8_scale_kernel); round to nearest to I8_FRAC fractional bits (|error| <= 2^-(I8_FRAC+1) of the row
            // scale, the clamp never bites), then I8_DIG balanced radix-256 digits: integer arithmetic only (i8digits.cuh)
            signed char dg[I8_DIG];
            i8_split(i8_fixed((lp[b] * ecol[(long long)s * mp_total + b]) * inv), dg);
#pragma unroll
            for (int i = 0; i < I8_DIG; ++i) d[i][b] = dg[i];
        }
#pragma unroll
        for (int i = 0; i < I8_DIG; ++i) {
            signed char* img = G + gbase + (((long long)blk * T + s) * (long long)I8_B_ALL);
            int off;
            if (layout == 0) {
                off = ((cc * (I8_DIG * 8) + (i * 8 + r1)) * 8 + r0) * 16;             // one 384-row operand
            } else {
                // cta_group::2: the N = 256 MMA takes rows 0-127 from the leader and 128-255 from the peer, the N = 128
                // MMA rows 0-63 / 64-127; each CTA's half is a 192-row operand [128 rows | 64 rows], halves back to back
                const int g = i * 64 + r;                                           // row of the 384-row operand
                const int h = g < I8_NA ? g / (I8_NA / 2) : (g - I8_NA) / (I8_NB / 2);
                const int lr = g < I8_NA ? g % (I8_NA / 2) : I8_NA / 2 + (g - I8_NA) % (I8_NB / 2);
                off = h * (I8_B_ALL / 2) + ((cc * (I8_DIG * 4) + (lr >> 3)) * 8 + (lr & 7)) * 16;
            }
            *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<const uint4*>(d[i]);
        }
    }
}

// ------------------------------------------------------------------------------------------ main kernel
__device__ __forceinline__ void i8_tmem_ld4(unsigned taddr, int (&v)[4]) {       // 4 consecutive accumulator columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void i8_tmem_ld1(unsigned taddr, int& v) {            // one accumulator column
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
// wait for the thread's TMEM loads; the destination registers are operands, so nothing that reads them moves above
__device__ __forceinline__ void i8_tmem_wait(int (&d)[I8_DIG][4]) {
#if GMS_I8_DIG == 7
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(d[0][0]), "+r"(d[0][1]), "+r"(d[0][2]), "+r"(d[0][3]), "+r"(d[1][0]), "+r"(d[1][1]), "+r"(d[1][2]), "+r"(d[1][3]),
                   "+r"(d[2][0]), "+r"(d[2][1]), "+r"(d[2][2]), "+r"(d[2][3]), "+r"(d[3][0]), "+r"(d[3][1]), "+r"(d[3][2]), "+r"(d[3][3]),
                   "+r"(d[4][0]), "+r"(d[4][1]), "+r"(d[4][2]), "+r"(d[4][3]), "+r"(d[5][0]), "+r"(d[5][1]), "+r"(d[5][2]), "+r"(d[5][3]),
                   "+r"(d[6][0]), "+r"(d[6][1]), "+r"(d[6][2]), "+r"(d[6][3])
                 :: "memory");
#else
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(d[0][0]), "+r"(d[0][1]), "+r"(d[0][2]), "+r"(d[0][3]), "+r"(d[1][0]), "+r"(d[1][1]), "+r"(d[1][2]), "+r"(d[1][3]),
                   "+r"(d[2][0]), "+r"(d[2][1]), "+r"(d[2][2]), "+r"(d[2][3]), "+r"(d[3][0]), "+r"(d[3][1]), "+r"(d[3][2]), "+r"(d[3][3]),
                   "+r"(d[4][0]), "+r"(d[4][1]), "+r"(d[4][2]), "+r"(d[4][3]), "+r"(d[5][0]), "+r"(d[5][1]), "+r"(d[5][2]), "+r"(d[5][3])
                 :: "memory");
#endif
}

// Loop order: trait s outer, 64-row tile inner; one unit = the accumulator of one (trait, tile). TT >= T is the compile-time
// size of the column Z[:, s] an epilogue thread keeps in registers (EXACT: TT == T, no run-time trait predicates; every T <= 10).
// CG2 = true: launched as clusters of 2 CTAs = two adjacent cross tiles forming ONE M = 256 tile (tcgen05 cta_group::2).
// Each CTA stages its own xi tile and only HALF of every digit tile; the leader CTA's MMA thread issues the pair's MMAs once
// both halves have landed (the peer relays its `full` phases to the leader), and its commits release the slot / publish the
// accumulators in BOTH CTAs. This halves the shared-memory traffic per MAC, which is what bounds the single-CTA kernel.
//
// The epilogue of a unit has two phases. DRAIN: tcgen05.ld of the warp's 32 rows x I8_DIG digit sums, recombined on the fly
// into 32 exact 64-bit integers W[j] that stay in registers, then the accumulator is released - the next unit's MMAs wait
// only for ~50 TMEM loads and ~300 integer instructions per warp (round 1 released after 7/8 of ALL the epilogue work of
// the unit: its MMA warp spent 36 % of its time waiting for the accumulator, ncu source page of profiles/r1_i8_kernel_ncu_final).
// MATH: conversion to FP64 and the T FMAs per row, overlapped with the next unit's MMAs. W[] needs compile-time register
// indices, so the rows are a fully unrolled sequence of warp-uniform predicated blocks - affordable only with ONE copy of it,
// hence the trait-outer loop for every T (the tile-outer loop of round 1 unrolled the epilogue T times to keep Z[T][T]).
template <int TT, bool EXACT, bool CG2>
__global__ void __launch_bounds__(I8_THREADS, 1) i8_contract_kernel(const I8Params p) {
    const int T = EXACT ? TT : p.T;          // EXACT: T == TT at compile time (every T <= 10 has its own instantiation)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stage0 = smem;
    constexpr int I8_STAGE_BYTES = I8Geo<CG2>::STAGE_BYTES, B_STAGE = I8Geo<CG2>::B_STAGE;
    const int I8_STAGES = p.stages;          // as many as fit next to the work list and the row data (i8_stages)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + I8_STAGES * I8_STAGE_BYTES);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * I8_MAX_STAGES + 2);
    int4* ttab = reinterpret_cast<int4*>(bars + 2 * I8_MAX_STAGES + 4);        // this CTA's range of the 64-row tile list (I8_TTAB entries)
    // row data of a unit, three buffers in turn: d_t[j] for T traits x 64 rows (FP64), then eb[j] = exponent field of scale_s[j] + 63 -
    // I8_FRAC - 1 (int) for the unit's trait s
    unsigned char* rowbuf = reinterpret_cast<unsigned char*>(ttab + I8_TTAB);
    const int rowbuf_bytes = T * 64 * 8 + 64 * 4;
    const unsigned full0 = i8_smem_u32(bars), empty0 = i8_smem_u32(bars + I8_MAX_STAGES);
    const unsigned accfull = i8_smem_u32(bars + 2 * I8_MAX_STAGES), accempty = i8_smem_u32(bars + 2 * I8_MAX_STAGES + 1);
    const unsigned crank = CG2 ? i8_cta_rank() : 0u;       // 0 = leader of the pair
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long X = blockIdx.x;
    const int split = blockIdx.y;
    const int jt_begin = p.split_begin[split], jt_end = p.split_begin[split + 1];
    // the CTA's range of the work list, once, in shared memory: (J | nkb << 16, padded index of the tile's first row, index of
    // its first (block, trait) digit image, 64-blocks of the chromosome) - no role chases jtiles -> chrom -> gbase through
    // global memory per unit. The unit's nkb k-blocks are J, J + 1, ... modulo the chromosome's block count (common.cuh)
    for (int i = threadIdx.x; i < jt_end - jt_begin; i += I8_THREADS) {
        const RowTile t = p.jtiles[jt_begin + i];
        const ChromDesc cd = p.chrom[t.c];
        const int nb = (cd.m + I8_KB - 1) / I8_KB;            // 64-blocks of the chromosome; row tile J owns i8_nassign(nb, J) of them
        const long long gblk = p.gbase[t.c] / I8_B_ALL + i8_blocks_before(nb, t.I) * p.T;      // < 2^31 (checked by the host plan)
        ttab[i] = make_int4(t.I | (i8_nassign(nb, t.I) << 16), cd.pad_off + 64 * t.I, (int)gblk, nb);
    }
    // every role walks the units (64-row tile, trait) in the same order: f(tile entry, s). GROUP-OUTER: the CTA's tile range is
    // cut into groups = runs of tiles of one chromosome (a tile with J == 0 opens a group; the list is ordered by (c, J)), and
    // the order is group, trait, tile. All T passes over a chromosome then re-read the same xi k-blocks (128 crosses x m_c bytes
    // per CTA: 35 MB over the 148 resident CTAs at cassava scale) and the chromosome's digit images from L2. With the trait
    // outermost (every trait pass walks the CTA's whole range, 218 MB of xi over the resident CTAs) ncu counted 14.0 GB of
    // DRAM reads per launch against 2.9 GB of operand images.
    // -DGMS_I8_TRAIT_OUTER (tools only, for A/B timing): one group = the whole range, i.e. trait, tile.
    auto group_end = [&](int g0, int n) {
#ifdef GMS_I8_TRAIT_OUTER
        (void)g0;
        return n;
#else
        int g1 = g0 + 1;
        while (g1 < n && (ttab[g1].x & 0xFFFF) != 0) ++g1;
        return g1;
#endif
    };
    auto for_units = [&](auto f) {
        const int n = jt_end - jt_begin;
        for (int g0 = 0; g0 < n;) {
            const int g1 = group_end(g0, n);
            for (int s = 0; s < T; ++s)
                for (int jt = g0; jt < g1; ++jt) f(ttab[jt], s);
            g0 = g1;
        }
    };

    if (threadIdx.x == 0) {
        // CG2 leader: a `full` phase needs its own data AND the peer's relay; accumulators are drained by both CTAs' epilogues
        for (int s = 0; s < I8_STAGES; ++s) { i8_mbar_init(full0 + 8 * s, (CG2 && crank == 0) ? 2 : 1); i8_mbar_init(empty0 + 8 * s, 1); }
        i8_mbar_init(accfull, 1);
        i8_mbar_init(accempty, CG2 ? 2 * I8_EPI_WARPS : I8_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == I8_EPI_WARPS) {
        if (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG2) i8_cluster_sync();           // the peer's barriers and TMEM must exist before anything touches them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    if (warp == I8_EPI_WARPS) {
        // ------------------------------------------------------------------ producer: the whole warp walks the loop (addresses and
        // counters stay in uniform registers), one elected lane issues the copies
        unsigned slot = 0, ph = 0;
        for_units([&](const int4 e, int s) {
            const int J = e.x & 0xFFFF, nkb = e.x >> 16, nb = e.w;
            const long long gblk = (long long)e.z;
            const signed char* abase = p.xi + (X * p.nkb_total + ((e.y - 64 * J) >> 6)) * (long long)I8_A_BYTES;       // k-block 0 of the chromosome
            const signed char* gb = p.G + (gblk + s) * (long long)I8_B_ALL + crank * B_STAGE;       // CG2: this CTA's half
            int ka = J;                                  // the chromosome's k-block of step kb: i8_block_k(nb, J, kb), stepped
            for (int kb = 0; kb < nkb; ++kb) {
                i8_mbar_wait(empty0 + 8 * slot, ph ^ 1u);
                if (v2_elect()) {
                    const unsigned dst = i8_smem_u32(stage0) + slot * I8_STAGE_BYTES, bar = full0 + 8 * slot;
                    if (I8_DBG(8 | 16)) {            // profiling: without the xi copy (8) / the digit copy (16)
                        i8_mbar_expect_tx(bar, (I8_DBG(8) ? 0 : I8_A_BYTES) + (I8_DBG(16) ? 0 : B_STAGE));
                        if (!I8_DBG(8)) i8_bulk_g2s(dst, abase + (long long)ka * I8_A_BYTES, I8_A_BYTES, bar);
                        if (!I8_DBG(16)) i8_bulk_g2s(dst + I8_A_BYTES, gb + (long long)kb * T * I8_B_ALL, B_STAGE, bar);
                    } else {
                        i8_mbar_expect_tx(bar, I8_STAGE_BYTES);
                        i8_bulk_g2s(dst, abase + (long long)ka * I8_A_BYTES, I8_A_BYTES, bar);
                        i8_bulk_g2s(dst + I8_A_BYTES, gb + (long long)kb * T * I8_B_ALL, B_STAGE, bar);
                    }
                }
                __syncwarp();
#ifndef GMS_I8_TRIANGLE
                if (++ka == nb) ka = 0;
#else
                ka = kb;                                 // triangle order: J, 0, 1, .. J - 1
#endif
                if (++slot == I8_STAGES) { slot = 0; ph ^= 1u; }
            }
        });
    } else if (warp == I8_EPI_WARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer (CG2: the leader; the peer relays)
        if (CG2 && crank != 0) {
            // peer of a pair: relay every completed `full` phase (own xi tile + own half of the digits) to the leader
            unsigned slot = 0, ph = 0;
            for_units([&](const int4 e, int) {
                const int nkb = e.x >> 16;
                for (int kb = 0; kb < nkb; ++kb) {
                    i8_mbar_wait(full0 + 8 * slot, ph);
                    if (v2_elect()) i8_mbar_arrive_remote(full0 + 8 * slot, 0);
                    __syncwarp();
                    if (++slot == I8_STAGES) { slot = 0; ph ^= 1u; }
                }
            });
        } else {
            // instruction descriptor: D = S32, A = B = signed 8 bit, K-major both; M = 128 (256 over a CTA pair)
            const unsigned idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)((CG2 ? 2 * I8_M : I8_M) >> 4) << 24);
            const unsigned idescA = idesc0 | ((unsigned)(I8_NA >> 3) << 17);           // N = 256: digits 0-3
            const unsigned idescB = idesc0 | ((unsigned)(I8_NB >> 3) << 17);           // N = 192 / 128: the remaining digits
            // rows of the digit operand held in THIS CTA's shared memory: all of them, or half of each MMA's (the pair's halves)
            constexpr unsigned B_ROWS = CG2 ? (I8_DIG * I8_N) / 2 : I8_DIG * I8_N;
            constexpr unsigned B_LBO = (B_ROWS / 8) * 128;                           // chunk stride
            constexpr unsigned B_SECOND = (CG2 ? I8_NA / 2 : I8_NA) / 8 * 128;       // byte offset of the second MMA's part
            // descriptors of slot 0, k32-step 0; everything else is an add on the 14-bit address field
            const unsigned long long da0 = i8_desc(i8_smem_u32(stage0), I8_M * 16, 128);
            const unsigned long long db0 = i8_desc(i8_smem_u32(stage0) + I8_A_BYTES, B_LBO, 128);
            unsigned slot = 0, ph = 0, u = 0;
            auto issue = [&](unsigned sl, int kb) {      // the MMAs of one 64-k stage (elected lane)
                const unsigned long long das = da0 + ((sl * I8_STAGE_BYTES) >> 4), dbs = db0 + ((sl * I8_STAGE_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < I8_KB / 32; ++kk) {
                    if (I8_DBG(2)) continue;
                    const unsigned long long da = das + ((kk * 2 * (I8_M * 16)) >> 4);
                    const unsigned long long dbA = dbs + ((kk * 2 * B_LBO) >> 4), dbB = dbs + ((B_SECOND + kk * 2 * B_LBO) >> 4);
                    if (CG2) {
                        i8_mma_cg2(tmem_base, da, dbA, idescA, (kb | kk) != 0);
                        if (!I8_DBG(4)) i8_mma_cg2(tmem_base + I8_NA, da, dbB, idescB, (kb | kk) != 0);
                    } else {
                        i8_mma(tmem_base, da, dbA, idescA, (kb | kk) != 0);
                        if (!I8_DBG(4)) i8_mma(tmem_base + I8_NA, da, dbB, idescB, (kb | kk) != 0);
                    }
                }
            };
            for_units([&](const int4 e, int) {
                const int nkb = e.x >> 16;
                i8_mbar_wait(accempty, (u & 1u) ^ 1u);                     // the epilogue(s) have drained the accumulators
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb) {
                    i8_mbar_wait(full0 + 8 * slot, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (v2_elect()) {
                        issue(slot, kb);
                        if (CG2) i8_commit_cg2(empty0 + 8 * slot, 3);     // frees the slot in BOTH CTAs when these MMAs retire
                        else i8_commit(empty0 + 8 * slot);
                        if (kb + 1 == nkb) {                               // accumulators of this (tile, trait) are complete
                            if (CG2) i8_commit_cg2(accfull, 3);
                            else i8_commit(accfull);
                        }
                    }
                    __syncwarp();
                    if (++slot == I8_STAGES) { slot = 0; ph ^= 1u; }
                }
                ++u;
            });
        }
    } else {
        // ------------------------------------------------------------------ epilogue: thread = one cross x one half of the rows
        const int quarter = warp & 3, half = warp >> 2;       // TMEM lane quarter is fixed by warp id % 4; `half`: rows 32 half .. + 31
        const int etid = threadIdx.x;                          // 0 .. 255
        const long long unit = X * I8_M + quarter * 32 + lane;
        const bool live = unit < p.n_units;
        int pa = 0, pb = 0;
        if (live) { pa = p.unit_a[unit]; pb = (p.mode == MASK_XI) ? p.unit_b[unit] : 0; }
        const unsigned tlane = tmem_base + ((unsigned)(quarter * 32) << 16) + half * I8_RPT;

        auto release = [&]() {                       // TMEM fully read: the next unit's MMAs may start
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) { if (CG2) i8_mbar_arrive_remote(accempty, 0); else i8_mbar_arrive(accempty); }
        };
        auto load4 = [&](int ch, int (&d)[I8_DIG][4]) {        // rows 4 ch .. 4 ch + 3 of this warp's half, all digit planes
#pragma unroll
            for (int i = 0; i < I8_DIG; ++i) i8_tmem_ld4(tlane + i * I8_N + ch * 4, d[i]);
        };
        // FP64 FMAs have a long latency and a column Z[:, s] has only T accumulation chains: NP partial sums per trait (rows
        // jr = k mod NP) keep >= 10 independent FMAs in flight per warp (measured: one chain per trait made the burst 3 x longer,
        // four per trait spill at T = 5)
#ifndef GMS_I8_NP5
#define GMS_I8_NP5 2
#endif
#ifndef GMS_I8_NP8
#define GMS_I8_NP8 2
#endif
        constexpr int NP = TT <= 5 ? GMS_I8_NP5 : (TT <= 8 ? GMS_I8_NP8 : 1);
        double zs[TT][NP];
#pragma unroll
        for (int t = 0; t < TT; ++t)
#pragma unroll
            for (int k = 0; k < NP; ++k) zs[t][k] = 0.0;
        // row data of unit u + 1 is fetched while unit u is processed: NPF values per thread of the (T + 1) x 64 block
        // [d_t[j] for T traits | scale_s[j]] plus the thread's four bit-plane words
        constexpr int NPF = (TT + 1 + 3) / 4;
        const int ntl = jt_end - jt_begin, nunits = T * ntl;
        double pf[NPF];
        unsigned pm[4] = {0u, 0u, 0u, 0u};
        auto pf_issue = [&](int s_, int jt_) {
            const long long prow = ttab[jt_].y;
#pragma unroll
            for (int k = 0; k < NPF; ++k) {
                const int i = etid + k * (32 * I8_EPI_WARPS);
                if (i < (T + 1) * 64) {
                    const int t = i >> 6, j = i & 63;
                    pf[k] = __ldg((t < T ? p.emat + (long long)t * p.mp_total : p.scale + (long long)s_ * p.mp_total) + prow + j);
                }
            }
            if (live) {
                const long long wi = (prow >> 5) + half;
                pm[0] = __ldg(p.planeA + (long long)pa * p.wpp + wi);
                pm[1] = __ldg(p.planeB + (long long)pa * p.wpp + wi);
                if (p.mode == MASK_XI) {
                    pm[2] = __ldg(p.planeA + (long long)pb * p.wpp + wi);
                    pm[3] = __ldg(p.planeB + (long long)pb * p.wpp + wi);
                }
            }
        };
        // ---- FP64 burst of one unit, the only FP64 instructions of the kernel: zs[t] += d_t[j] v_j for the unit's 32 rows, v_j as
        // bit patterns in V[], d_t[j] from the unit's row buffer. `rows` == 0: no unit of this warp has a mask entry in the tile
        // (a warp of padding units). Writes the finished column Z[:, s] out when the unit was the last tile of trait s.
        long long V[I8_RPT];
        auto fma_rows = [&](const double* erow, int jbeg, int jend) {       // rows [jbeg, jend) of the unit whose values are in V[]
#pragma unroll
            for (int j0 = jbeg; j0 < jend; j0 += 2) {
                const double v0 = __longlong_as_double(V[j0]), v1 = __longlong_as_double(V[j0 + 1]);
#pragma unroll
                for (int t = 0; t < TT; ++t)
                    if (t < T) {
                        const double2 e = *reinterpret_cast<const double2*>(erow + t * 64 + j0);      // d_t of two rows
                        zs[t][j0 % NP] = fma(e.x, v0, zs[t][j0 % NP]);
                        zs[t][(j0 + 1) % NP] = fma(e.y, v1, zs[t][(j0 + 1) % NP]);
                    }
            }
        };
        // the two row halves of a cross go to two partial slabs (summed by finalize_sym); a slab entry belongs to this thread
        // alone, so the groups after the first add to it with a plain read-modify-write, in a fixed order
        double* const zout = p.Zpart + ((long long)(split * 2 + half) * p.n_units + unit) * (T * T);
        constexpr int NPRE = TT <= 5 ? TT : 1;       // T <= 5: the old slab values are loaded before the FMAs of the burst (registers to spare)
        auto finish_trait = [&](int s_, bool first_group, const double (&zold)[NPRE]) {      // the column Z[:, s_] of one group of tiles is complete
            if (live) {
#pragma unroll
                for (int t = 0; t < TT; ++t)
                    if (t < T) {
                        double z = zs[t][0];
#pragma unroll
                        for (int k = 1; k < NP; ++k) z += zs[t][k];
                        if (NPRE == TT) zout[t * T + s_] = zold[t < NPRE ? t : 0] + z;
                        else zout[t * T + s_] = first_group ? z : zout[t * T + s_] + z;
                    }
            }
#pragma unroll
            for (int t = 0; t < TT; ++t)
#pragma unroll
                for (int k = 0; k < NP; ++k) zs[t][k] = 0.0;
        };
        auto burst = [&](const double* erow, unsigned rows, int last_of_trait, int s_) {       // last_of_trait: 0 no, 1 yes, 2 yes + first group
            double zold[NPRE];                       // what the slab holds from the earlier groups: loaded before the FMAs, used after
#pragma unroll
            for (int t = 0; t < NPRE; ++t) zold[t] = (NPRE == TT && last_of_trait == 1 && live && t < T) ? zout[t * T + s_] : 0.0;
            if (rows != 0u && !I8_DBG(64)) fma_rows(erow, 0, I8_RPT);
            if (last_of_trait) finish_trait(s_, last_of_trait == 2, zold);
        };
        if (nunits > 0) pf_issue(0, 0);
        else if (live) {                     // a range without tiles (the host plan makes none): its slab still has to read as zero
            for (int i = 0; i < T * T; ++i) zout[i] = 0.0;
        }
        // GMS_I8_LATE (default): the FP64 burst of unit u runs at the top of iteration u + 1, after `accfull` - the window in
        // which the tensor pipe has nothing to do anyway (the MMAs of unit u + 1 are done, those of u + 2 wait for the drain):
        // FP64 instructions next to running tcgen05 MMAs stall them (tools/tc_contention_probe.cu), here they cannot.
        // Measured at cassava scale: 37.0 ms against 39.1 ms with the burst right after the conversion (-DGMS_I8_LATE=0).
        // The row data of a unit then lives for two iterations: three row buffers. (Interleaving the burst with the drain loop,
        // chunk by chunk in the same registers, was no faster: 37.5 ms - the drain loses its double-buffered TMEM loads.)
#ifndef GMS_I8_LATE
#define GMS_I8_LATE 1
#endif
        constexpr bool LATE = GMS_I8_LATE && TT <= 10;           // T > 10: V[] next to a long column Z[:, s] would spill
        int s = 0, jl = 0;                                       // unit u = (trait s, tile jt_begin + jl), jl in the group [g0, g1)
        int g0 = 0, g1 = ntl > 0 ? group_end(0, ntl) : 0;
        unsigned rows_prev = 0u;
        int last_prev = 0;
        int s_prev = 0;
        for (int u = 0; u < nunits; ++u) {
            const unsigned acc_it = (unsigned)u;
            unsigned char* rb = rowbuf + (acc_it % 3u) * rowbuf_bytes;
            double* erow_all = reinterpret_cast<double*>(rb);                 // d_t[j]: [t][64]
            int* eb_all = reinterpret_cast<int*>(rb + T * 64 * 8);            // eb[j] of trait s: [64]
            // commit the prefetched values of this unit ...
#pragma unroll
            for (int k = 0; k < NPF; ++k) {
                const int i = etid + k * (32 * I8_EPI_WARPS);
                if (i < T * 64) erow_all[i] = pf[k];
                else if (i < (T + 1) * 64)   // the row scale is a power of two >= 2^-900: its exponent field, biased for i8_cvt_bits
                    eb_all[i & 63] = (__double2hiint(pf[k]) >> 20) + (63 - I8_FRAC - 1);
            }
            // this thread's mask entries of the tile's rows (one word of the bit-planes): nz = non-zero, ng = negative
            unsigned nz = 0u, ng = 0u;
            if (live) {
                if (p.mode == MASK_XI) {
                    nz = (pm[0] ^ pm[1]) & (pm[2] ^ pm[3]);
                    ng = ((~pm[0] & pm[1]) ^ (~pm[2] & pm[3])) & nz;
                } else {
                    nz = pm[0] ^ pm[1];
                    ng = (p.mode == MASK_DELTA) ? (~pm[0] & pm[1]) : 0u;
                }
            }
            // ... and start the loads of the next one (for_units order: group, trait, tile)
            int s_n = s, jl_n = jl + 1, g0_n = g0, g1_n = g1;
            if (jl_n == g1) {
                if (++s_n == T) { s_n = 0; g0_n = g1; g1_n = g1 < ntl ? group_end(g1, ntl) : g1; }
                jl_n = g0_n;
            }
            const int last_now = (jl + 1 == g1) ? (g0 == 0 ? 2 : 1) : 0;       // this unit ends the pass of trait s over the group
            if (u + 1 < nunits) pf_issue(s_n, jl_n);
            const unsigned rows = __reduce_or_sync(0xffffffffu, nz);       // warp-uniform: rows some unit of the warp needs
            asm volatile("bar.sync 1, %0;" ::"n"(32 * I8_EPI_WARPS) : "memory");
            const double* erow = erow_all + half * I8_RPT;       // erow[t * 64 + jr]
            const int* ebrow = eb_all + half * I8_RPT;            // ebrow[jr]
            i8_mbar_wait(accfull, acc_it & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const double* erow_prev = reinterpret_cast<const double*>(rowbuf + ((acc_it + 2u) % 3u) * rowbuf_bytes) + half * I8_RPT;
            if (LATE && u > 0 && !I8_DBG(1 | 32)) burst(erow_prev, rows_prev, last_prev, s_prev);
            if (I8_DBG(1)) {
                release();
            } else {
                // ---- drain: TMEM -> 32 exact row sums in registers; loads of chunk c + 1 in flight while chunk c is recombined.
                // The one part of the epilogue no MMA can overlap with while a unit's accumulator fills TMEM
                {
                    int dA[I8_DIG][4], dB[I8_DIG][4];
                    auto comb4 = [&](int ch, const int (&d)[I8_DIG][4]) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            int sv[I8_DIG];
#pragma unroll
                            for (int i = 0; i < I8_DIG; ++i) sv[i] = d[i][q];
                            V[ch * 4 + q] = i8_combine(sv);
                        }
                    };
                    constexpr int NCH = I8_RPT / 4;          // 8 chunks of 4 rows
                    load4(0, dA);
                    i8_tmem_wait(dA);
#pragma unroll
                    for (int cp = 0; cp < NCH / 2; ++cp) {
                        load4(2 * cp + 1, dB);
                        comb4(2 * cp, dA);
                        i8_tmem_wait(dB);
                        if (cp + 1 < NCH / 2) load4(2 * cp + 2, dA);
                        comb4(2 * cp + 1, dB);
                        if (cp + 1 < NCH / 2) i8_tmem_wait(dA);
                    }
                }
                release();
                if (!I8_DBG(32)) {
                    // ---- convert (integer instructions only: overlaps the next unit's MMAs for free):
                    // V[j] <- bit pattern of v_j = xi_x[j] W[j] 2^(E_j - I8_FRAC)
#pragma unroll
                    for (int jr = 0; jr < I8_RPT; ++jr) {
                        unsigned bh, bl;
                        i8_cvt_bits(V[jr], ebrow[jr], (nz >> jr) & 1u, (ng >> jr) & 1u, bh, bl);
                        V[jr] = (long long)(((unsigned long long)bh << 32) | bl);
                    }
                    // Every conversion is complete before the first FP64 instruction, and the FP64 instructions form ONE dense
                    // burst per unit: FMAs scattered over integer code cost the tensor pipe ~4-5 cycles per warp instruction.
                    // The empty asm pins the values for the compiler.
#pragma unroll
                    for (int jr = 0; jr < I8_RPT; ++jr) asm volatile("" : "+l"(V[jr]));
                    if (!LATE) burst(erow, rows, last_now, s);
                }
            }
            rows_prev = rows; last_prev = last_now; s_prev = s;
            s = s_n; jl = jl_n; g0 = g0_n; g1 = g1_n;
        }
        if (LATE && nunits > 0 && !I8_DBG(1 | 32))
            burst(reinterpret_cast<const double*>(rowbuf + ((unsigned)(nunits - 1) % 3u) * rowbuf_bytes) + half * I8_RPT, rows_prev, last_prev, s_prev);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG2) i8_cluster_sync();           // the pair's MMAs read both CTAs' shared memory and write both TMEMs
    if (warp == I8_EPI_WARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int i8_npart(int T) { (void)T; return 2; }                 // partial slabs per split: one per row half
int i8_max_tiles_per_split() { return I8_TTAB; }
long long i8_block_bytes() { return I8_B_ALL; }            // digit image of one (64 x 64 block, trait)
static size_t i8_fixed_smem(int T) {
    const size_t rowbuf = (size_t)T * 64 * 8 + 64 * 4;          // per buffer (see the epilogue)
    return (2 * I8_MAX_STAGES + 4) * 8 + (size_t)I8_TTAB * 16 + 3 * rowbuf;          // three row buffers: see GMS_I8_LATE
}
// pipeline stages: as many as fit in the 227 KiB a CTA may use (the kernel is bound by the latency of its operand stream)
int i8_stages(int T, bool cg2) {
    const size_t stage = cg2 ? I8Geo<true>::STAGE_BYTES : I8Geo<false>::STAGE_BYTES;
    const int n = (int)((232448 - i8_fixed_smem(T)) / stage);
    return n < 2 ? 2 : (n > I8_MAX_STAGES ? I8_MAX_STAGES : n);
}
size_t i8_smem_bytes(int T, bool cg2) {
    return (size_t)i8_stages(T, cg2) * (cg2 ? I8Geo<true>::STAGE_BYTES : I8Geo<false>::STAGE_BYTES) + i8_fixed_smem(T);
}

cudaError_t launch_i8_build_xi(const unsigned* planeA, const unsigned* planeB, long long wpp, const int* sire,
                               const int* dam, long long n_units, int mode, long long nkb_total, signed char* xi,
                               cudaStream_t st) {
    const long long npad = (n_units + 2 * I8_M - 1) / (2 * I8_M) * (2 * I8_M);
    const long long total = npad * nkb_total;           // one thread per (unit, 64-k block)
    if (total == 0) return cudaSuccess;
    i8_build_xi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(planeA, planeB, wpp, sire, dam, n_units, mode,
                                                                         nkb_total, xi);
    return cudaGetLastError();
}

cudaError_t launch_i8_build_digits(const double* tiles, const ChromDesc* chrom_d, int c_begin, int c_end, int C,
                                   const int* blkpref_d, int nblks, const double* emat, long long mp_total, int T,
                                   double* scale, double* smax, const long long* gbase_d, signed char* G, int layout,
                                   cudaStream_t st) {
    if (nblks == 0) return cudaSuccess;
    // `scale` arrives zeroed (cudaMemsetAsync in api.cu); it holds the row maxima until i8_scale_kernel
    i8_rowmax_kernel<<<nblks, 256, 0, st>>>(tiles, chrom_d, c_begin, c_end, blkpref_d, emat, mp_total, T, reinterpret_cast<unsigned long long*>(scale));
    cudaError_t e = launch_i8_smax(scale, chrom_d, c_begin, c_end, C, mp_total, T, smax, st);     // verify.cu: per-chromosome largest scale
    if (e != cudaSuccess) return e;
    i8_scale_kernel<<<(unsigned)(((long long)T * mp_total + 255) / 256), 256, 0, st>>>(scale, (long long)T * mp_total);
    i8_build_digits_kernel<<<nblks, 256, 0, st>>>(tiles, chrom_d, c_begin, c_end, blkpref_d, emat, mp_total, T, scale, gbase_d, G, layout);
    return cudaGetLastError();
}

cudaError_t launch_i8_contract(const I8Params& p, int ntiles, int nsplit, bool pair, cudaStream_t st) {
    if (ntiles == 0) return cudaSuccess;
    const size_t smem = i8_smem_bytes(p.T, pair);
    I8Params pp = p;
    pp.stages = i8_stages(p.T, pair);
    cudaError_t e = cudaSuccess;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pair ? (ntiles + 1) / 2 * 2 : ntiles, nsplit, 1);
    cfg.blockDim = dim3(I8_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pair ? 1 : 0;
#define GMS_I8_ONE(TV, EX, PR)                                                                                                  \
    {                                                                                                                           \
        e = cudaFuncSetAttribute(i8_contract_kernel<TV, EX, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        if (e != cudaSuccess) return e;                                                                                         \
        e = cudaLaunchKernelEx(&cfg, i8_contract_kernel<TV, EX, PR>, pp);                                                       \
        if (e != cudaSuccess) return e;                                                                                         \
    }
#define GMS_I8(TV, EX) { if (pair) GMS_I8_ONE(TV, EX, true) else GMS_I8_ONE(TV, EX, false) }
    switch (p.T) {
        case 1: GMS_I8(1, true) break;
        case 2: GMS_I8(2, true) break;
        case 3: GMS_I8(3, true) break;
        case 4: GMS_I8(4, true) break;
        case 5: GMS_I8(5, true) break;
        case 6: GMS_I8(6, true) break;
        case 7: GMS_I8(7, true) break;
        case 8: GMS_I8(8, true) break;
        case 9: GMS_I8(9, true) break;
        case 10: GMS_I8(10, true) break;
        default:
            if (p.T <= 12) GMS_I8(12, false)
            else if (p.T <= 16) GMS_I8(16, false)
            else if (p.T <= 24) GMS_I8(24, false)
            else if (p.T <= 32) GMS_I8(32, false)
            else return cudaErrorInvalidValue;
    }
#undef GMS_I8
#undef GMS_I8_ONE
    return cudaGetLastError();
}

}  // namespace gms
