// Shared constants and device-side descriptors of the cross-variance engine (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gms {

// ---- tiling of the contraction  C[rows j, cols (unit,trait)] = L'[j,k] * W[k,(unit,trait)] ----
constexpr int BM = 128;                 // rows of one row-tile I (SNP j)
constexpr int BK = 32;                  // k per pipeline stage == one 32-bit haplotype word
constexpr int BN = 80;                  // columns per unit-tile: NU = BN / T units x T traits
constexpr int CHUNK = BM * BK;          // doubles per A chunk (32 KiB), fragment-major image
constexpr int BCHUNK = BK * BN;         // doubles per generated B chunk (20 KiB)
constexpr int MATH_WARPS = 8;           // 4 (M) x 2 (N); warp tile 32 x 40
constexpr int PROD_WARPS = 2;           // operand producers (bulk copy + on-the-fly W generation)
constexpr int NTHREADS = 32 * (MATH_WARPS + PROD_WARPS);
constexpr int MAX_T = 32;

enum MaskMode : int { MASK_DELTA = 0, MASK_HET = 1, MASK_XI = 2 };

// One chromosome of the padded SNP axis. Chromosome c occupies padded indices
// [pad_off, pad_off + mp), mp = roundup(m, 128); padded SNPs have zero effects and zero bits.
struct ChromDesc {
    long long tile_off;  // offset (doubles) of the chromosome's first chunk in a tile image
    int m;               // real SNPs
    int mp;              // padded SNPs (multiple of BM)
    int snp_off;         // first SNP in the caller's (unpadded) order
    int pad_off;         // first padded index (multiple of BM)
};

struct RowTile { int c; int I; };

// Tile image layout (built by recomb.cu, consumed by contract.cu via cp.async.bulk):
// per chromosome only the lower block triangle is stored: row-tile I owns chunks kc = 0 .. 4(I+1)-1
// (k = 32 kc .. 32 kc + 31), chunk index within the chromosome = 2 I (I+1) + kc. The diagonal
// 128 x 128 block is stored times 0.5, so that  x' M y = Z[t,s] + Z[s,t]  with Z from the lower
// block triangle only (M symmetric).
// Inside a chunk (4096 doubles) the order is the mma.m16n8k4 A-fragment order:
//   e = ((q * 8 + mb) * 32 + lane) * 2 + hi,  q = k4-step (0..7), mb = 16-row block (0..7)
//   row = 16 mb + (lane >> 2) + 8 hi,  k = 4 q + (lane & 3)
__host__ __device__ inline long long chunks_before_rowtile(int I) { return 2LL * I * (I + 1); }
__host__ __device__ inline long long chunks_of_chrom(int nI) { return 2LL * nI * (nI + 1); }

// ---- int8 path: which 64 x 64 blocks of a chromosome's symmetric matrix a 64-row tile J contracts (i8path.cu) ----
// x' M y = sum over unordered block pairs {J, K}; a pair may be evaluated in EITHER orientation, because
// B_KJ[t,s] + B_KJ[s,t] = B_JK[s,t] + B_JK[t,s] for symmetric M and the table is X = Z + Z'. Instead of the lower block
// triangle (row tile J owns K = 0 .. J: 1 to nb k-blocks per accumulator, and the epilogue of a short unit is longer than the
// next unit's MMAs) the pairs are dealt round-robin: with h = nb / 2, row tile J owns K = (J + i) mod nb, i = 0 .. n_J - 1,
//   nb odd:  n_J = h + 1 for every J;   nb even: n_J = h + 1 for J < h (they take the antipodal pair {J, J + h}), else h.
// Every unordered pair is owned exactly once (circular distance d <= h has one orientation with (K - J) mod nb = d, the
// antipodal distance of an even nb has two and goes to the smaller J); the block count stays nb (nb + 1) / 2 and every
// accumulator now sums (nb + 1) / 2 k-blocks, give or take one. i = 0 is the diagonal block, stored times 0.5.
// -DGMS_I8_TRIANGLE (tools only, for A/B timing) restores the lower block triangle: row tile J owns K = 0 .. J.
__host__ __device__ inline long long i8_blocks_of_chrom(int nb) { return (long long)nb * (nb + 1) / 2; }
#ifndef GMS_I8_TRIANGLE
__host__ __device__ inline int i8_nassign(int nb, int J) { const int h = nb >> 1; return (nb & 1) ? h + 1 : (J < h ? h + 1 : h); }
__host__ __device__ inline int i8_block_k(int nb, int J, int i) { const int k = J + i; return k >= nb ? k - nb : k; }   // i-th k-block of row tile J
__host__ __device__ inline long long i8_blocks_before(int nb, int J) {      // blocks owned by row tiles 0 .. J - 1
    const long long h = nb >> 1;
    if (nb & 1) return (long long)J * (h + 1);
    return J < h ? (long long)J * (h + 1) : h * (h + 1) + (J - h) * h;
}
__host__ __device__ inline void i8_block_decode(int nb, long long blk, int& J, int& i) {      // inverse of i8_blocks_before(J) + i
    const long long h = nb >> 1;
    if ((nb & 1) || blk < h * (h + 1)) { J = (int)(blk / (h + 1)); i = (int)(blk - (long long)J * (h + 1)); }
    else { const long long b = blk - h * (h + 1); J = (int)(h + b / h); i = (int)(b % h); }
}
#else
__host__ __device__ inline int i8_nassign(int nb, int J) { (void)nb; return J + 1; }
__host__ __device__ inline int i8_block_k(int nb, int J, int i) { (void)nb; return i == 0 ? J : i - 1; }      // diagonal first, then 0 .. J - 1
__host__ __device__ inline long long i8_blocks_before(int nb, int J) { (void)nb; return (long long)J * (J + 1) / 2; }
__host__ __device__ inline void i8_block_decode(int nb, long long blk, int& J, int& i) {
    (void)nb;
    J = 0;
    while ((long long)(J + 1) * (J + 2) / 2 <= blk) ++J;
    i = (int)(blk - (long long)J * (J + 1) / 2);
}
#endif

struct ContractParams {
    const double* tiles;         // R or RoR tile image
    const ChromDesc* chrom;
    const RowTile* rowtiles;     // work list, ordered by (c, I)
    const int* split_begin;      // [nsplit + 1] ranges of rowtiles
    const double* emat;          // T x mp_total effects on the padded axis
    long long mp_total;
    const unsigned* planeA;      // [P][wpp] HapA bits
    const unsigned* planeB;      // [P][wpp] HapB bits
    long long wpp;               // words per parent = mp_total / 32
    const int* unit_a;           // parent (DELTA/HET) or sire (XI) of each unit
    const int* unit_b;           // dam (XI)
    long long n_units;           // units (or, with n_units_dev, the capacity of the unit lists and of Z)
    const int* n_units_dev;      // NULL, or the device-side unit count (<= n_units is enforced): the grid is then fixed and
                                 // every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    int T;
    int NU;                      // units per tile = BN / T
    int mode;
    int stages;
    double* Z;                   // [nsplit][n_units][T*T]: Z[t][s] = sum_j e_t[j] mask[j] (L' W_s)[j]
};

size_t contract_smem_bytes(int T, int stages);
int contract_pick_stages(int T, size_t smem_limit);
cudaError_t launch_contract(const ContractParams& p, int ntiles, int nsplit, cudaStream_t st);

}  // namespace gms
