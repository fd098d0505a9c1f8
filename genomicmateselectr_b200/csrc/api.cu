// C ABI (include/gms_b200.h) and the engine handle: owns the HBM-resident LD-decay tile images,
// the bit-packed haplotypes, the per-parent tables and the work plan, and drives the kernels on one stream.
//
// One gms_compute = two phases:
//   parents : everything derived from the effects (padded effect matrices, scan records, int8 digit images) and the
//             per-parent tables Q_P / P_P (indexed by parent id) for the parents of the cross list that are not cached;
//             with gms_set_parent_shard only this rank's slice of them, packed into an exchange buffer that the host
//             all-gathers (NCCL, or cudaMemcpyPeer inside gms_predict_sharded)
//   crosses : the per-cross dominance term, Nsegsnps, assembly of the result slab
// In GMS_MODE_INT8_TENSOR every table is followed by the a-posteriori error bound of verify.cu; units that fail it are
// re-evaluated by the FP64 contraction in the same stream (no host round trip); if more fail than the lists hold the
// whole call is redone in FP64 when the host next synchronises.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gms_b200.h"
#include "common.cuh"
#include "kernels.cuh"

using namespace gms;

namespace {

std::mutex g_err_mu;
std::string g_create_err = "";

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t ensure(size_t n) {
        if (n <= cap && p) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        if (n == 0) n = 1;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() {}     // freed explicitly in gms_destroy (the right device must be current)
};

struct Plan {       // one contraction launch
    int ntiles = 0, nsplit = 1;
    std::vector<int> split_begin;
};

// one effect class: 0 = add (R, delta masks) -> VarA, 1 = dom (RoR, het / xi masks) -> VarD, 2 = asub (R, delta) -> VarBV
struct EffClass {
    DevBuf<double> raw, emat, S, scale, smax, table;
    DevBuf<signed char> G;
    void release() { raw.release(); emat.release(); S.release(); scale.release(); smax.release(); table.release(); G.release(); }
};

constexpr int FB_NSPLIT = 8;          // row-tile ranges of the FP64 re-evaluation launches
constexpr int FB_CAP_MAX = 16384;     // units per table that can be re-evaluated in-stream
constexpr int NEV = 8;

}  // namespace

struct gms_handle {
    std::recursive_mutex mu;           // entry points are serialised per handle
    int device = 0;
    int num_sms = 148;
    size_t smem_limit = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;

    // geometry
    int C = 0;
    std::vector<ChromDesc> chrom;
    DevBuf<ChromDesc> chrom_d;
    long long m = 0, mp_total = 0, wpp = 0, tile_doubles = 0;
    int c_begin = 0, c_end = 0;
    bool have_tiles = false, have_haps = false, staged = false, parents_done = false, computed = false, settled = false;
    DevBuf<double> R, R2;
    unsigned long long epoch = 0;      // bumped whenever tiles / haplotypes / shard / mode change: invalidates the cache

    // haplotypes
    long long P = 0;
    DevBuf<unsigned> planeA, planeB;

    // work list for the chromosome shard
    std::vector<RowTile> rowtiles;
    DevBuf<RowTile> rowtiles_d;
    double lower_elems = 0;   // padded elements in the stored lower block triangles of the shard
    long long S2 = 0;

    // staged problem
    int model = 0, T = 0, npairs = 0, ncomp = 0;
    long long N = 0;                   // crosses of THIS handle (its range of the staged list)
    EffClass cls[3];
    DevBuf<int> sire_d, dam_d, pair1_d, pair2_d, split_par_d, split_x_d, split_fb_d, flag_d;
    DevBuf<int> sire_s_d, dam_s_d, perm_d;      // int8 path: the crosses of every pass ordered by sire, and where each one goes in X
    DevBuf<double> Z, Zfb, X, out_d, gebv_d, mean_d, cm_add, cm_dom, cm_raw, tidy_d;
    DevBuf<int> nseg_d;
    Plan plan_par, plan_x, plan_fb;
    int stages = 0;

    // parents: tables are indexed by parent id; `have` = rows valid for the cached effects
    std::vector<unsigned char> have;
    std::vector<int> todo;             // parents whose table rows this call computes (all ranks: the same list)
    DevBuf<int> todo_d;
    long long todo_lo = 0, todo_n = 0, todo_L = 0;   // this rank's slice of todo; slice length of the exchange layout
    int par_index = 0, par_parts = 1;  // gms_set_parent_shard
    DevBuf<double> xbuf;               // exchange buffer [parts * L][ncls][T*T]
    bool cache_on = true, cache_hit = false, prep_needed = true;
    unsigned long long cache_key = 0, cross_key = 0, raw_key = 0, overflow_key = 0;
    bool overflow_valid = false;
    bool cache_valid = false;
    long long nP = 0;                  // distinct parents of the staged list

    // int8 tensor path
    std::vector<RowTile> jtiles;
    DevBuf<RowTile> jtiles_d;
    std::vector<long long> gbase;
    DevBuf<long long> gbase_d;
    DevBuf<signed char> xi_d, xi_het_d, xi_delta_d;
    DevBuf<int> split_i8_d, split_i8_par_d, split_i8_tail_d, i8_blkpref_d;
    int i8_nblks = 0;
    Plan plan_i8, plan_i8_par, plan_i8_tail;
    long long i8_chunk = 0;            // crosses per pass of the per-cross int8 kernel (bounds the xi image)
    bool use_i8 = false;
    bool i8_pair = true;               // tcgen05 cta_group::2: CTA pairs form one M = 256 tile (GMS_I8_CG2=0: one CTA per 128 crosses)
    bool i8_sort = true;               // per-cross passes run in sire order (GMS_I8_SORT=0: list order), so that the crosses of a warp share a parent
    bool xi_image_valid = false;       // the per-cross xi image in xi_d belongs to the staged cross list (single pass only)
    double tau = 0x1p-31;              // int8 error-bound tolerance (gms_set_int8_tolerance); <= 0: no verification
    int fb_cap = 0;
    DevBuf<int> fb_counter, fb_a, fb_b, fb_row, ver_n;
    DevBuf<double> ver_b;
    int n_counters = 0;
    std::vector<int> counters_h;
    int i8_debug = 0;

    // mode
    int mode = GMS_MODE_AUTO;          // requested
    int eff_mode = GMS_MODE_DENSE;     // what the staged problem will run (AUTO resolved)
    long long i8_gbytes = 0;
    bool scan_ok = false;
    DevBuf<double> scan_ab;            // [4][mp]: a(0.02), b(0.02), a(0.04), b(0.04)
    DevBuf<int> split_scan_par_d, split_scan_x_d;
    std::vector<int> scan_split_par, scan_split_x;

    cudaEvent_t ev[NEV] = {};
    gms_stats st{};
};

namespace {

int fail(gms_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    else { std::lock_guard<std::mutex> lk(g_err_mu); g_create_err = msg; }
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(h, e__ == cudaErrorMemoryAllocation ? GMS_ENOMEM : GMS_ECUDA,         \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                 \
    } while (0)
#define LOCK(h) std::lock_guard<std::recursive_mutex> lock__((h)->mu)

int bind(gms_handle* h) {
    CU(cudaSetDevice(h->device));
    return GMS_OK;
}

void invalidate(gms_handle* h) {
    ++h->epoch;
    h->cache_valid = false;
    h->overflow_valid = false;
    h->xi_image_valid = false;
    h->staged = h->parents_done = h->computed = false;
}

// validate first, commit after: a failing call leaves the previous geometry intact
int set_geometry(gms_handle* h, int C, const int64_t* m_c) {
    if (C <= 0 || !m_c) return fail(h, GMS_EINVAL, "need C >= 1 chromosomes and m_c");
    std::vector<ChromDesc> chrom((size_t)C);
    long long snp = 0, pad = 0, tiles = 0;
    for (int c = 0; c < C; ++c) {
        if (m_c[c] <= 0 || m_c[c] > (1LL << 30)) return fail(h, GMS_EINVAL, "m_c[c] must be in 1..2^30");
        ChromDesc& cd = chrom[c];
        cd.m = (int)m_c[c];
        cd.mp = (int)((m_c[c] + BM - 1) / BM * BM);
        cd.snp_off = (int)snp;
        cd.pad_off = (int)pad;
        cd.tile_off = tiles;
        snp += cd.m;
        pad += cd.mp;
        tiles += chunks_of_chrom(cd.mp / BM) * CHUNK;
        if (pad > (1LL << 31) - 1) return fail(h, GMS_EINVAL, "padded SNP axis exceeds 2^31");
    }
    invalidate(h);
    h->have_tiles = h->have_haps = false;      // from here on the old state is gone
    CU(h->chrom_d.ensure(C));
    CU(h->R.ensure((size_t)tiles));
    CU(h->R2.ensure((size_t)tiles));
    h->chrom.swap(chrom);
    h->C = C;
    h->m = snp;
    h->mp_total = pad;
    h->wpp = pad / 32;
    h->tile_doubles = tiles;
    h->c_begin = 0;
    h->c_end = C;
    CU(cudaMemcpyAsync(h->chrom_d.p, h->chrom.data(), sizeof(ChromDesc) * C, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

int build_worklist(gms_handle* h) {
    h->rowtiles.clear();
    h->lower_elems = 0;
    h->S2 = 0;
    for (int c = h->c_begin; c < h->c_end; ++c) {
        const int nI = h->chrom[c].mp / BM;
        for (int I = 0; I < nI; ++I) {
            h->rowtiles.push_back(RowTile{c, I});
            // what contract_kernel really issues: 16-row blocks with a real row x k-chunks with a real k
            const int rows = h->chrom[c].m - BM * I;
            const int nact = rows >= BM ? 8 : (rows + 15) / 16;
            const int nkc = std::min(4 * (I + 1), (h->chrom[c].m + BK - 1) / BK);
            h->lower_elems += 16.0 * nact * BK * nkc;
        }
        h->S2 += (long long)h->chrom[c].m * h->chrom[c].m;
    }
    CU(h->rowtiles_d.ensure(h->rowtiles.size()));
    CU(cudaMemcpyAsync(h->rowtiles_d.p, h->rowtiles.data(), sizeof(RowTile) * h->rowtiles.size(),
                       cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

// equal-work contiguous ranges of a (c, I)-ordered tile list (weight I + 1); ranges no longer than `cap` tiles if cap > 0
// (weights: k-blocks per tile when they are not I + 1 - the int8 path deals the blocks evenly, common.cuh)
void cut_ranges(const std::vector<RowTile>& tiles, int ns, int cap, Plan& pl, const std::vector<int>* weights = nullptr) {
    const int n = (int)tiles.size();
    ns = std::max(1, std::min(ns, std::max(1, n)));
    if (cap > 0) ns = std::max(ns, (n + cap - 1) / cap);
    auto wt = [&](int i) { return weights ? (double)(*weights)[i] : (double)(tiles[i].I + 1); };
    double total = 0;
    for (int i = 0; i < n; ++i) total += wt(i);
    for (;; ++ns) {
        pl.nsplit = ns;
        pl.split_begin.assign(ns + 1, 0);
        double acc = 0;
        int sp = 1;
        for (int i = 0; i < n && sp < ns; ++i) {
            acc += wt(i);
            // close range sp-1 after tile i once its share is reached, keeping >= 1 tile for the rest - and never an empty
            // range (a heavy tile can reach several shares at once): a CTA without tiles would leave its partial slab unwritten
            while (sp < ns && acc >= total * sp / ns && (n - (i + 1)) >= (ns - sp) && i + 1 > pl.split_begin[sp - 1]) pl.split_begin[sp++] = i + 1;
        }
        for (; sp < ns; ++sp) pl.split_begin[sp] = std::max(pl.split_begin[sp - 1], n - (ns - sp));
        pl.split_begin[ns] = n;
        int longest = 0;
        for (int i = 1; i <= ns; ++i) {
            pl.split_begin[i] = std::max(pl.split_begin[i], pl.split_begin[i - 1]);
            longest = std::max(longest, pl.split_begin[i] - pl.split_begin[i - 1]);
        }
        if (cap <= 0 || longest <= cap || ns >= n) break;
    }
}

// number of ranges: fill >= 2 waves of SMs for few unit tiles, minimise the last-wave loss for many
int pick_nsplit(int ctas_per_split, int sms) {
    if (ctas_per_split <= 0) return 1;
    if (ctas_per_split < 2 * sms) return (2 * sms + ctas_per_split - 1) / ctas_per_split;
    double best = 1e30;
    int ns = 1;
    for (int cand = 1; cand <= 8; ++cand) {
        const double ctas = (double)ctas_per_split * cand, waves = std::ceil(ctas / sms);
        const double cost = waves * sms / ctas * (1.0 + 0.01 * (cand - 1));
        if (cost < best - 1e-12) { best = cost; ns = cand; }
    }
    return ns;
}

void make_plan(const gms_handle* h, long long n_units, int NU, Plan& pl) {
    pl.ntiles = (int)((n_units + NU - 1) / NU);
    cut_ranges(h->rowtiles, pick_nsplit(pl.ntiles, h->num_sms), 0, pl);
}

int upload_plan(gms_handle* h, const Plan& pl, DevBuf<int>& dev) {
    CU(dev.ensure(pl.split_begin.size()));
    CU(cudaMemcpyAsync(dev.p, pl.split_begin.data(), sizeof(int) * pl.split_begin.size(), cudaMemcpyHostToDevice, h->stream));
    return GMS_OK;
}

// FP64 contraction of `n_units` units into table rows. row_index: unit u -> row (NULL: u). n_dev: device-side count.
int run_contract(gms_handle* h, const double* tiles, const double* emat, int mode, const int* ua, const int* ub,
                 long long n_units, const Plan& pl, const int* split_d, double* Zbuf, double* Xout, const int* row_index,
                 const int* n_dev = nullptr, cudaEvent_t e_before = nullptr, cudaEvent_t e_after = nullptr) {
    if (n_units == 0) return GMS_OK;
    ContractParams p{};
    p.tiles = tiles;
    p.chrom = h->chrom_d.p;
    p.rowtiles = h->rowtiles_d.p;
    p.split_begin = split_d;
    p.emat = emat;
    p.mp_total = h->mp_total;
    p.planeA = h->planeA.p;
    p.planeB = h->planeB.p;
    p.wpp = h->wpp;
    p.unit_a = ua;
    p.unit_b = ub;
    p.n_units = n_units;
    p.n_units_dev = n_dev;
    p.T = h->T;
    p.NU = BN / h->T;
    p.mode = mode;
    p.stages = h->stages;
    p.Z = Zbuf;
    int gx = pl.ntiles;
    if (n_dev) gx = std::max(1, std::min(pl.ntiles, 2 * h->num_sms / pl.nsplit));   // fixed grid: CTAs walk the live tiles
    if (e_before) CU(cudaEventRecord(e_before, h->stream));
    CU(launch_contract(p, gx, pl.nsplit, h->stream));
    if (e_after) CU(cudaEventRecord(e_after, h->stream));
    CU(launch_finalize_sym(Zbuf, pl.nsplit, n_units, h->T, Xout, h->stream, true, row_index, n_dev, n_units));
    h->st.last_launches += 2;
    return GMS_OK;
}

// chromosome ranges for the scan path: enough (unit x range) threads to fill the GPU, bounded partial buffer
void make_scan_split(const gms_handle* h, long long n_units, std::vector<int>& sb) {
    const int nchr = h->c_end - h->c_begin;
    long long want = n_units > 0 ? (600000 + n_units - 1) / n_units : 1;
    const long long cap = std::max<long long>(1, (1LL << 30) / std::max<long long>(1, n_units * h->T * h->T * 8));
    int ns = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(want, cap), nchr));
    sb.assign(ns + 1, h->c_begin);
    double total = 0, acc = 0;
    for (int c = h->c_begin; c < h->c_end; ++c) total += h->chrom[c].m;
    int sp = 1;
    for (int c = h->c_begin; c < h->c_end && sp < ns; ++c) {
        acc += h->chrom[c].m;
        while (sp < ns && acc >= total * sp / ns && (h->c_end - (c + 1)) >= (ns - sp)) sb[sp++] = c + 1;
    }
    for (; sp < ns; ++sp) sb[sp] = std::max(sb[sp - 1], h->c_end - (ns - sp));
    sb[ns] = h->c_end;
    for (int i = 1; i <= ns; ++i) sb[i] = std::max(sb[i], sb[i - 1]);
}

int run_scan(gms_handle* h, const double* S, int mode, const int* ua, const int* ub, long long n_units,
             const std::vector<int>& sb, const int* split_d, double* Xout, const int* row_index,
             cudaEvent_t e_before = nullptr, cudaEvent_t e_after = nullptr) {
    if (n_units == 0) return GMS_OK;
    ScanParams p{};
    p.chrom = h->chrom_d.p;
    p.split_begin = split_d;
    p.S = S;
    p.planeA = h->planeA.p;
    p.planeB = h->planeB.p;
    p.wpp = h->wpp;
    p.unit_a = ua;
    p.unit_b = ub;
    p.n_units = n_units;
    p.T = h->T;
    p.mode = mode;
    p.Xpart = h->Z.p;
    const int ns = (int)sb.size() - 1;
    if (e_before) CU(cudaEventRecord(e_before, h->stream));
    CU(launch_scan(p, ns, h->stream));
    if (e_after) CU(cudaEventRecord(e_after, h->stream));
    CU(launch_finalize_sym(h->Z.p, ns, n_units, h->T, Xout, h->stream, false, row_index));
    h->st.last_launches += 2;
    return GMS_OK;
}

// int8 tcgen05 contraction of `n_units` units (their xi image in `xi`) into table rows
int run_i8(gms_handle* h, const EffClass& cl, const signed char* xi, int mode, const int* ua, const int* ub, long long n_units,
           const Plan& pl, const int* split_d, double* Xout, const int* row_index, cudaEvent_t e0 = nullptr,
           cudaEvent_t e1 = nullptr) {
    if (n_units == 0) return GMS_OK;
    I8Params ip{};
    ip.chrom = h->chrom_d.p; ip.jtiles = h->jtiles_d.p; ip.split_begin = split_d;
    ip.xi = xi; ip.nkb_total = h->mp_total / 64; ip.G = cl.G.p; ip.gbase = h->gbase_d.p;
    ip.scale = cl.scale.p; ip.emat = cl.emat.p; ip.mp_total = h->mp_total;
    ip.planeA = h->planeA.p; ip.planeB = h->planeB.p; ip.wpp = h->wpp;
    ip.unit_a = ua; ip.unit_b = ub; ip.n_units = n_units; ip.T = h->T; ip.mode = mode;
    ip.Zpart = h->Z.p;
    ip.debug = h->i8_debug;
    if (e0) CU(cudaEventRecord(e0, h->stream));
    CU(launch_i8_contract(ip, pl.ntiles, pl.nsplit, h->i8_pair, h->stream));
    if (e1) CU(cudaEventRecord(e1, h->stream));
    CU(launch_finalize_sym(h->Z.p, i8_npart(h->T) * pl.nsplit, n_units, h->T, Xout, h->stream, true, row_index));
    h->st.last_launches += 2;
    return GMS_OK;
}

// the accuracy contract of the int8 path (verify.cu): bound every unit of the table just written, list the units that
// fail, re-evaluate those with the FP64 contraction - all on the stream, the host only reads the counter afterwards
int verify_and_repair(gms_handle* h, const EffClass& cl, const double* tiles, int mode, const int* ua, const int* ub,
                      long long n_units, double* table, bool rows_by_parent, int counter_idx) {
    if (n_units == 0 || h->tau <= 0.0) return GMS_OK;
    VerifyParams v{};
    v.chrom = h->chrom_d.p; v.c_begin = h->c_begin; v.c_end = h->c_end; v.C = h->C;
    v.planeA = h->planeA.p; v.planeB = h->planeB.p; v.wpp = h->wpp;
    v.unit_a = ua; v.unit_b = ub; v.n_units = n_units; v.mode = mode;
    v.emat = cl.emat.p; v.mp_total = h->mp_total; v.T = h->T; v.smax = cl.smax.p;
    v.table = table; v.rows_by_parent = rows_by_parent ? 1 : 0; v.tau = h->tau;
    v.counter = h->fb_counter.p + counter_idx; v.cap = h->fb_cap;
    v.list_a = h->fb_a.p; v.list_b = h->fb_b.p; v.list_row = h->fb_row.p;
    v.part_b = h->ver_b.p; v.part_n = h->ver_n.p;
    CU(launch_i8_verify(v, h->stream));
    h->st.last_launches += 2;
    return run_contract(h, tiles, cl.emat.p, mode, h->fb_a.p, h->fb_b.p, h->fb_cap, h->plan_fb, h->split_fb_d.p, h->Zfb.p, table,
                        h->fb_row.p, h->fb_counter.p + counter_idx);
}

// 64-bit hash of a block of doubles + finiteness check in the same pass (4 independent lanes)
unsigned long long hash_doubles(const double* e, size_t n, unsigned long long seed, long long& bad) {
    const unsigned long long K = 0x9E3779B97F4A7C15ull;
    unsigned long long a[4] = {seed ^ K, seed + 0x632BE59BD9B4E019ull, seed ^ 0xD6E8FEB86659FD93ull, seed + 0xA0761D6478BD642Full};
    unsigned long long inf_acc = 0;      // stays 0 while every exponent field is below 0x7FF
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(e);
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const unsigned long long x = w[i + l];
            inf_acc |= ((x >> 52) & 0x7FF) == 0x7FF;
            a[l] = (a[l] ^ x) * K;
            a[l] ^= a[l] >> 29;
        }
    }
    for (; i < n; ++i) {
        const unsigned long long x = w[i];
        inf_acc |= ((x >> 52) & 0x7FF) == 0x7FF;
        a[0] = (a[0] ^ x) * K;
        a[0] ^= a[0] >> 29;
    }
    if (inf_acc && bad < 0)
        for (size_t j = 0; j < n; ++j)
            if (!std::isfinite(e[j])) { bad = (long long)j; break; }
    unsigned long long r = a[0];
    for (int l = 1; l < 4; ++l) r = (r ^ a[l]) * K + (r >> 31);
    return r ^ (unsigned long long)n;
}
unsigned long long hash_doubles_mt(const double* e, size_t n, unsigned long long seed, long long& bad) {
    const size_t chunk = 1u << 21;
    if (n <= 2 * chunk) return hash_doubles(e, n, seed, bad);
    const size_t nch = (n + chunk - 1) / chunk;
    std::vector<unsigned long long> hs(nch);
    std::vector<long long> bads(nch, -1);
    const unsigned nth = std::max(1u, std::min(8u, std::min((unsigned)nch, std::thread::hardware_concurrency())));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nth; ++t)
        th.emplace_back([&, t]() {
            for (size_t c = t; c < nch; c += nth) {
                const size_t lo = c * chunk, len = std::min(chunk, n - lo);
                hs[c] = hash_doubles(e + lo, len, seed + c, bads[c]);
                if (bads[c] >= 0) bads[c] += (long long)lo;
            }
        });
    for (auto& t : th) t.join();
    unsigned long long r = seed;
    for (size_t c = 0; c < nch; ++c) {
        r = (r ^ hs[c]) * 0x9E3779B97F4A7C15ull + (r >> 29);
        if (bads[c] >= 0 && bad < 0) bad = bads[c];
    }
    return r;
}
unsigned long long hash_ints(const int32_t* a, size_t n, unsigned long long seed) {
    unsigned long long r = seed ^ 0x2545F4914F6CDD1Dull;
    for (size_t i = 0; i < n; ++i) { r = (r ^ (unsigned)a[i]) * 0x9E3779B97F4A7C15ull; r ^= r >> 32; }
    return r ^ n;
}

int make_i8_plan(gms_handle* h, long long units, Plan& pl, DevBuf<int>& dev) {
    pl.ntiles = (int)((units + 127) / 128);
    std::vector<int> w(h->jtiles.size());
    for (size_t i = 0; i < w.size(); ++i) w[i] = i8_nassign((h->chrom[h->jtiles[i].c].m + 63) / 64, h->jtiles[i].I);
    cut_ranges(h->jtiles, pick_nsplit(pl.ntiles, h->num_sms), i8_max_tiles_per_split(), pl, &w);
    return upload_plan(h, pl, dev);
}

// ---- stage: inputs to HBM, mode resolution, plans, buffers. The handle processes crosses [xlo, xhi) of the list; the
// parents whose table rows must exist are those of the WHOLE list, so that all ranks of a parent-sharded run agree.
int stage_impl(gms_handle* h, int model, int T, const double* add, const double* dom, const double* asub, int64_t Nall,
               const int32_t* sire, const int32_t* dam, int64_t xlo, int64_t xhi) {
    if (!h->have_tiles || !h->have_haps) return fail(h, GMS_ESTATE, "map and haplotypes must be set before predict");
    if (model < GMS_MODEL_A || model > GMS_MODEL_DIRDOM) return fail(h, GMS_EINVAL, "unknown model");
    if (T < 1 || T > MAX_T) return fail(h, GMS_EINVAL, "T must be in 1..32");
    if (!add) return fail(h, GMS_EINVAL, "add effects are NULL");
    if (model >= GMS_MODEL_AD && !dom) return fail(h, GMS_EINVAL, "dominance effects are required for AD / DirDom");
    if (model == GMS_MODEL_DIRDOM && !asub) return fail(h, GMS_EINVAL, "allele-substitution effects are required for DirDom");
    if (Nall < 0 || (Nall > 0 && (!sire || !dam))) return fail(h, GMS_EINVAL, "bad cross list");
    if (xlo < 0 || xhi < xlo || xhi > Nall) return fail(h, GMS_EINVAL, "bad cross range");
    const long long N = xhi - xlo;
    if (N > (1LL << 31) - 1) return fail(h, GMS_EINVAL, "too many crosses in one call");
    const int ncls = model + 1;
    const double* host_eff[3] = {add, dom, asub};
    // one host pass over the effects: finiteness (NA / NaN / Inf would propagate to every variance in the reference; here they
    // are an error, like haplotype values outside {0, 1}) and the hash that keys the per-handle cache
    unsigned long long key = 0xC0FFEEull + (unsigned long long)model * 131 + (unsigned long long)T;
    for (int k = 0; k < ncls; ++k) {
        long long bad = -1;
        key = hash_doubles_mt(host_eff[k], (size_t)h->m * T, key + k, bad);
        if (bad >= 0) {
            char buf[160];
            snprintf(buf, sizeof buf, "non-finite marker effect (%s effects, SNP %lld, trait %lld)",
                     k == 0 ? "additive" : (k == 1 ? "dominance" : "allele-substitution"), bad % h->m, bad / h->m);
            return fail(h, GMS_EINVAL, buf);
        }
    }
    // parents of the whole list
    std::vector<unsigned char> need((size_t)h->P, 0);
    for (long long x = 0; x < Nall; ++x) {
        const int s = sire[x], d = dam[x];
        if (s < 0 || s >= h->P || d < 0 || d >= h->P) {
            char buf[128];
            snprintf(buf, sizeof buf, "cross %lld: parent index out of range (P = %lld)", (long long)x, h->P);
            return fail(h, GMS_EINVAL, buf);
        }
        need[s] = need[d] = 1;
    }
    if (int rc = bind(h)) return rc;
    h->staged = h->parents_done = h->computed = false;
    const bool geometry_changed = h->model != model || h->T != T;
    h->model = model;
    h->T = T;
    h->N = N;
    h->ncomp = ncls;
    h->npairs = T * (T + 1) / 2;
    const int TT = T * T;
    h->stages = contract_pick_stages(T, h->smem_limit);
    if (h->stages < 2) return fail(h, GMS_ECUDA, "not enough shared memory per block for the contraction kernel");

    // resolve the evaluation mode (from the sizes of the whole problem, so that it does not depend on what is cached)
    const int NU = BN / T;
    const long long npad = (N + 255) / 256 * 256;
    h->nP = 0;
    for (long long p = 0; p < h->P; ++p) h->nP += need[p];
    const int parts = h->par_parts;
    const long long ppad_max = ((h->nP + parts - 1) / parts + 255) / 256 * 256;
    double dig_bytes = 0;
    for (int c = h->c_begin; c < h->c_end; ++c) {
        const double nJ = (h->chrom[c].m + 63) / 64;
        dig_bytes += nJ * (nJ + 1) / 2 * T * (double)i8_block_bytes();
    }
    // int8 path: the parents' mask images (2 ppad x mp bytes) and up to 3 digit-plane sets must fit comfortably
    const bool i8_ok = Nall > 0 && (double)(2 * ppad_max) * (double)h->mp_total < 16e9 && dig_bytes * ncls < 60e9;
    const int prev_mode = h->eff_mode;
    h->eff_mode = h->mode;
    // AUTO: the int8 tensor path, unless it does not fit or the same effects already overflowed its error-bound lists
    if (h->mode == GMS_MODE_AUTO) h->eff_mode = (i8_ok && !(h->overflow_valid && key == h->overflow_key)) ? GMS_MODE_INT8_TENSOR : GMS_MODE_DENSE;
    const unsigned long long rawkey = key;
    if (h->eff_mode == GMS_MODE_INT8_TENSOR && !i8_ok) h->eff_mode = GMS_MODE_DENSE;
    h->use_i8 = h->eff_mode == GMS_MODE_INT8_TENSOR;
    // the per-cross pass is chunked so that the xi image stays below ~12 GB
    long long chunk = (long long)(12e9 / (double)h->mp_total) / 256 * 256;
    if (const char* ev = getenv("GMS_I8_CHUNK")) chunk = (atoll(ev) + 255) / 256 * 256;      // test hook
    chunk = std::max<long long>(256, std::min<long long>(chunk, std::max<long long>(256, npad)));
    h->i8_chunk = chunk;

    // the cache: same effects, model, T, resolved mode and geometry epoch -> the operand images in HBM and the table rows
    // computed so far stay valid; only parents that have no row yet are computed
    key = (key ^ h->epoch) * 0x9E3779B97F4A7C15ull + (unsigned long long)h->eff_mode;
    h->cache_hit = h->cache_on && h->cache_valid && key == h->cache_key && !geometry_changed && prev_mode == h->eff_mode;
    h->raw_key = rawkey;
    if (!h->cache_hit) {
        h->have.assign((size_t)h->P, 0);
        h->cache_valid = false;
        h->cache_key = key;
    }
    h->prep_needed = !h->cache_hit;
    h->todo.clear();
    for (long long p = 0; p < h->P; ++p)
        if (need[p] && !h->have[p]) h->todo.push_back((int)p);
    const long long ntodo = (long long)h->todo.size();
    h->todo_L = (ntodo + parts - 1) / parts;
    h->todo_lo = std::min<long long>(ntodo, h->par_index * h->todo_L);
    h->todo_n = std::min<long long>(ntodo, h->todo_lo + h->todo_L) - h->todo_lo;
    const long long nun = h->todo_n;                     // per-parent units of THIS rank
    const long long ppad = (nun + 255) / 256 * 256;

    // crosses of this rank
    const unsigned long long ckey = hash_ints(sire + xlo, (size_t)N, hash_ints(dam + xlo, (size_t)N, h->epoch));
    if (ckey != h->cross_key) h->xi_image_valid = false;
    h->cross_key = ckey;
    CU(h->sire_d.ensure((size_t)N)); CU(h->dam_d.ensure((size_t)N));
    if (N) {
        CU(cudaMemcpyAsync(h->sire_d.p, sire + xlo, sizeof(int) * N, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->dam_d.p, dam + xlo, sizeof(int) * N, cudaMemcpyHostToDevice, h->stream));
    }
    CU(h->todo_d.ensure((size_t)std::max<long long>(1, ntodo)));
    if (ntodo) CU(cudaMemcpyAsync(h->todo_d.p, h->todo.data(), sizeof(int) * ntodo, cudaMemcpyHostToDevice, h->stream));
    std::vector<int32_t> p1(h->npairs), p2(h->npairs);
    gms_trait_pairs(T, p1.data(), p2.data());
    CU(h->pair1_d.ensure(h->npairs)); CU(h->pair2_d.ensure(h->npairs));
    CU(cudaMemcpyAsync(h->pair1_d.p, p1.data(), sizeof(int) * h->npairs, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->pair2_d.p, p2.data(), sizeof(int) * h->npairs, cudaMemcpyHostToDevice, h->stream));

    // raw effects to HBM; everything derived from them (padded layout, scan records, digit planes) is built on the device
    // inside gms_compute, i.e. inside the timed region of a benchmark (skipped on a cache hit: nothing changed)
    for (int k = 0; k < ncls; ++k) {
        EffClass& cl = h->cls[k];
        const size_t n = (size_t)h->m * T;
        CU(cl.raw.ensure(n));
        CU(cl.emat.ensure((size_t)h->mp_total * T));
        CU(cl.table.ensure((size_t)h->P * TT));
        if (h->prep_needed) CU(cudaMemcpyAsync(cl.raw.p, host_eff[k], n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    if (parts > 1) CU(h->xbuf.ensure((size_t)std::max<long long>(1, h->todo_L * parts) * ncls * TT));

    size_t zneed = 0;
    if (h->eff_mode == GMS_MODE_MAP_SCAN) {
        const size_t sn = (size_t)h->mp_total * 3 * T;
        for (int k = 0; k < ncls; ++k) CU(h->cls[k].S.ensure(sn));
        make_scan_split(h, nun, h->scan_split_par);
        make_scan_split(h, model >= GMS_MODEL_AD ? N : 0, h->scan_split_x);
        CU(h->split_scan_par_d.ensure(h->scan_split_par.size()));
        CU(h->split_scan_x_d.ensure(h->scan_split_x.size()));
        CU(cudaMemcpyAsync(h->split_scan_par_d.p, h->scan_split_par.data(), sizeof(int) * h->scan_split_par.size(), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->split_scan_x_d.p, h->scan_split_x.data(), sizeof(int) * h->scan_split_x.size(), cudaMemcpyHostToDevice, h->stream));
        zneed = std::max<size_t>((h->scan_split_par.size() - 1) * (size_t)nun * TT,
                                 model >= GMS_MODEL_AD ? (h->scan_split_x.size() - 1) * (size_t)N * TT : (size_t)0);
    }
    if (h->eff_mode != GMS_MODE_MAP_SCAN) {
        // FP64 contraction plans: the dense mode itself, and the re-evaluation of units that fail the int8 error bound
        make_plan(h, nun, NU, h->plan_par);
        make_plan(h, model >= GMS_MODEL_AD ? N : 0, NU, h->plan_x);
        if (int rc = upload_plan(h, h->plan_par, h->split_par_d)) return rc;
        if (int rc = upload_plan(h, h->plan_x, h->split_x_d)) return rc;
        if (!h->use_i8)
            zneed = std::max<size_t>((size_t)h->plan_par.nsplit * (size_t)nun * TT,
                                     model >= GMS_MODEL_AD ? (size_t)h->plan_x.nsplit * (size_t)N * TT : (size_t)0);
    }
    if (h->use_i8) {
        // 64-row tiles (each owns about half of the chromosome's 64-blocks, common.cuh) + digit image offsets for the chromosome shard
        h->jtiles.clear();
        h->gbase.assign(h->C, 0);
        long long gb = 0;
        std::vector<int> blkpref(h->C + 1, 0);
        int blks = 0;
        for (int c = 0; c < h->C; ++c) {
            blkpref[c] = blks;
            if (c < h->c_begin || c >= h->c_end) continue;
            const int nJ = (h->chrom[c].m + 63) / 64;
            h->gbase[c] = gb;
            gb += (long long)nJ * (nJ + 1) / 2 * T * i8_block_bytes();
            blks += nJ * (nJ + 1) / 2;
            for (int J = 0; J < nJ; ++J) h->jtiles.push_back(RowTile{c, J});
        }
        blkpref[h->C] = blks;
        h->i8_nblks = blks;
        h->i8_gbytes = gb;
        CU(h->i8_blkpref_d.ensure(h->C + 1));
        CU(cudaMemcpyAsync(h->i8_blkpref_d.p, blkpref.data(), sizeof(int) * (h->C + 1), cudaMemcpyHostToDevice, h->stream));
        CU(h->jtiles_d.ensure(h->jtiles.size()));
        CU(h->gbase_d.ensure(h->C));
        CU(cudaMemcpyAsync(h->jtiles_d.p, h->jtiles.data(), sizeof(RowTile) * h->jtiles.size(), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->gbase_d.p, h->gbase.data(), sizeof(long long) * h->C, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));     // blkpref is a local
        // digit planes: (R, add) [and (R, asub)] for the additive tables, (RoR, dom) for the dominance terms
        for (int k = 0; k < ncls; ++k) {
            CU(h->cls[k].G.ensure((size_t)gb));
            CU(h->cls[k].scale.ensure((size_t)h->mp_total * T));
            CU(h->cls[k].smax.ensure((size_t)h->C * T));
        }
        // mask images: parents (delta, het), crosses (xi)
        CU(h->xi_delta_d.ensure((size_t)ppad * h->mp_total));
        if (model >= GMS_MODEL_AD) {
            CU(h->xi_het_d.ensure((size_t)ppad * h->mp_total));
            const size_t xin = (size_t)std::min<long long>(chunk, std::max<long long>(256, npad)) * h->mp_total;
            if (xin > h->xi_d.cap) h->xi_image_valid = false;
            CU(h->xi_d.ensure(xin));
        }
        if (model >= GMS_MODEL_AD && N > 0 && h->i8_sort) {
            // every pass of the per-cross kernel in sire order (stable: list order within a sire), so that the 32 crosses of an
            // epilogue warp share one parent as far as the list allows and their non-zero xi rows coincide (i8path.cu: SKIP).
            // xi = delta_sire o delta_dam is symmetric in the pair, so the key could be either parent; the sire it is.
            std::vector<int32_t> ss((size_t)N), ds((size_t)N), pm((size_t)N);
            std::vector<long long> cnt;
            std::vector<int32_t> idx;
            for (long long off = 0; off < N; off += chunk) {
                const long long n = std::min<long long>(chunk, N - off);
                const int32_t* sv = sire + xlo + off;
                const int32_t* dv = dam + xlo + off;
                if (h->P <= 4 * n + 4096) {            // counting sort
                    cnt.assign((size_t)h->P + 1, 0);
                    for (long long r = 0; r < n; ++r) ++cnt[(size_t)sv[r] + 1];
                    for (long long q = 0; q < h->P; ++q) cnt[q + 1] += cnt[q];
                    for (long long r = 0; r < n; ++r) {
                        const long long pos = off + cnt[sv[r]]++;
                        ss[pos] = sv[r]; ds[pos] = dv[r]; pm[pos] = (int32_t)r;
                    }
                } else {
                    idx.resize((size_t)n);
                    for (long long r = 0; r < n; ++r) idx[r] = (int32_t)r;
                    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return sv[a] < sv[b]; });
                    for (long long r = 0; r < n; ++r) { ss[off + r] = sv[idx[r]]; ds[off + r] = dv[idx[r]]; pm[off + r] = idx[r]; }
                }
            }
            CU(h->sire_s_d.ensure((size_t)N)); CU(h->dam_s_d.ensure((size_t)N)); CU(h->perm_d.ensure((size_t)N));
            CU(cudaMemcpyAsync(h->sire_s_d.p, ss.data(), sizeof(int) * N, cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(h->dam_s_d.p, ds.data(), sizeof(int) * N, cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(h->perm_d.p, pm.data(), sizeof(int) * N, cudaMemcpyHostToDevice, h->stream));
            CU(cudaStreamSynchronize(h->stream));     // the host vectors are locals
        }
        const long long n_full = model >= GMS_MODEL_AD ? std::min<long long>(N, chunk) : 0;
        const long long n_tail = model >= GMS_MODEL_AD && N > chunk ? N % chunk : 0;
        if (int rc = make_i8_plan(h, n_full, h->plan_i8, h->split_i8_d)) return rc;
        if (int rc = make_i8_plan(h, n_tail, h->plan_i8_tail, h->split_i8_tail_d)) return rc;
        if (int rc = make_i8_plan(h, nun, h->plan_i8_par, h->split_i8_par_d)) return rc;
        const size_t npart = (size_t)i8_npart(T);
        zneed = std::max<size_t>(std::max<size_t>(npart * h->plan_i8.nsplit * (size_t)n_full * TT, npart * h->plan_i8_tail.nsplit * (size_t)n_tail * TT),
                                 npart * h->plan_i8_par.nsplit * (size_t)nun * TT);
        // FP64 re-evaluation of units that fail the error bound: fixed grid, FB_NSPLIT row-tile ranges, bounded lists
        const long long most = std::max<long long>(nun, n_full);
        h->fb_cap = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(FB_CAP_MAX, most), (1LL << 28) / ((long long)FB_NSPLIT * TT * 8)));
        h->plan_fb.ntiles = (h->fb_cap + NU - 1) / NU;
        cut_ranges(h->rowtiles, FB_NSPLIT, 0, h->plan_fb);
        if (int rc = upload_plan(h, h->plan_fb, h->split_fb_d)) return rc;
        CU(h->Zfb.ensure((size_t)h->plan_fb.nsplit * h->fb_cap * TT));
        CU(h->fb_a.ensure(h->fb_cap)); CU(h->fb_b.ensure(h->fb_cap)); CU(h->fb_row.ensure(h->fb_cap));
        CU(h->ver_b.ensure(i8_verify_scratch_doubles(most, h->c_end - h->c_begin, T)));
        CU(h->ver_n.ensure((size_t)most * (h->c_end - h->c_begin)));
        const long long npass = model >= GMS_MODEL_AD && N > 0 ? (N + chunk - 1) / chunk : 0;
        h->n_counters = 3 + (int)npass;
        CU(h->fb_counter.ensure(h->n_counters));
    } else {
        h->n_counters = 0;
    }
    CU(h->Z.ensure(zneed));
    if (model >= GMS_MODEL_AD) CU(h->X.ensure((size_t)N * TT));
    CU(h->out_d.ensure((size_t)N * h->npairs * h->ncomp));
    CU(h->nseg_d.ensure((size_t)N));
    // the host vectors above go out of scope: make sure the copies are done
    CU(cudaStreamSynchronize(h->stream));
    h->staged = true;
    return GMS_OK;
}

const double* class_tiles(const gms_handle* h, int k) { return k == 1 ? h->R2.p : h->R.p; }
int class_mask(int k) { return k == 1 ? MASK_HET : MASK_DELTA; }

// everything derived from the effects, on the device
int prep_effects(gms_handle* h) {
    const int T = h->T, ncls = h->ncomp;
    for (int k = 0; k < ncls; ++k)
        CU(launch_scatter_effects(h->cls[k].raw.p, h->m, T, h->chrom_d.p, h->C, h->cls[k].emat.p, h->mp_total, h->stream));
    h->st.last_launches += ncls;
    if (h->eff_mode == GMS_MODE_MAP_SCAN) {
        const double* ab = h->scan_ab.p;
        for (int k = 0; k < ncls; ++k) {
            const double* a = k == 1 ? ab + 2 * h->mp_total : ab;
            CU(launch_scan_prepare(h->cls[k].emat.p, a, a + h->mp_total, h->mp_total, T, h->cls[k].S.p, h->stream));
        }
        h->st.last_launches += ncls;
    } else if (h->use_i8) {
        for (int k = 0; k < ncls; ++k) {
            EffClass& cl = h->cls[k];
            CU(cudaMemsetAsync(cl.scale.p, 0, sizeof(double) * h->mp_total * T, h->stream));
            CU(cudaMemsetAsync(cl.smax.p, 0, sizeof(double) * h->C * T, h->stream));
            CU(launch_i8_build_digits(class_tiles(h, k), h->chrom_d.p, h->c_begin, h->c_end, h->C, h->i8_blkpref_d.p, h->i8_nblks,
                                      cl.emat.p, h->mp_total, T, cl.scale.p, cl.smax.p, h->gbase_d.p, cl.G.p, h->i8_pair ? 1 : 0, h->stream));
            h->st.last_launches += 4;
        }
    }
    return GMS_OK;
}

int compute_parents_impl(gms_handle* h) {
    if (!h->staged) return fail(h, GMS_ESTATE, "gms_stage_inputs has not been called");
    if (int rc = bind(h)) return rc;
    h->st.last_launches = 0;
    h->computed = false;
    CU(cudaEventRecord(h->ev[0], h->stream));
    if (h->n_counters) CU(cudaMemsetAsync(h->fb_counter.p, 0, sizeof(int) * h->n_counters, h->stream));
    if (h->prep_needed) if (int rc = prep_effects(h)) return rc;
    const int ncls = h->ncomp, TT = h->T * h->T;
    const long long nun = h->todo_n;
    const int* units = h->todo_d.p + h->todo_lo;          // unit = parent id = table row
    if (nun > 0) {
        if (h->eff_mode == GMS_MODE_MAP_SCAN) {
            for (int k = 0; k < ncls; ++k)
                if (int rc = run_scan(h, h->cls[k].S.p, class_mask(k), units, nullptr, nun, h->scan_split_par,
                                      h->split_scan_par_d.p, h->cls[k].table.p, units)) return rc;
        } else if (h->use_i8) {
            CU(launch_i8_build_xi(h->planeA.p, h->planeB.p, h->wpp, units, nullptr, nun, MASK_DELTA, h->mp_total / 64, h->xi_delta_d.p, h->stream));
            if (ncls >= 2)
                CU(launch_i8_build_xi(h->planeA.p, h->planeB.p, h->wpp, units, nullptr, nun, MASK_HET, h->mp_total / 64, h->xi_het_d.p, h->stream));
            h->st.last_launches += 1 + (ncls >= 2);
            for (int k = 0; k < ncls; ++k) {
                const EffClass& cl = h->cls[k];
                if (int rc = run_i8(h, cl, k == 1 ? h->xi_het_d.p : h->xi_delta_d.p, class_mask(k), units, nullptr, nun, h->plan_i8_par,
                                    h->split_i8_par_d.p, h->cls[k].table.p, units)) return rc;
                if (int rc = verify_and_repair(h, cl, class_tiles(h, k), class_mask(k), units, nullptr, nun, h->cls[k].table.p, true, k)) return rc;
            }
        } else {
            // Q_P = (a o delta_P)' R (a o delta_P), P_P = (d o het_P)' RoR (d o het_P)   
            for (int k = 0; k < ncls; ++k)
                if (int rc = run_contract(h, class_tiles(h, k), h->cls[k].emat.p, class_mask(k), units, nullptr, nun, h->plan_par,
                                          h->split_par_d.p, h->Z.p, h->cls[k].table.p, units)) return rc;
        }
    }
    if (h->par_parts > 1 && !h->todo.empty()) {
        // this rank's rows -> its slice of the exchange buffer [parts * L][ncls][T*T]; the host all-gathers the slices
        const double* tabs[3] = {h->cls[0].table.p, h->cls[1].table.p, h->cls[2].table.p};
        CU(launch_pack_rows(tabs, ncls, TT, units, nun, h->xbuf.p + (size_t)h->par_index * h->todo_L * ncls * TT, h->stream));
        h->st.last_launches += 1;
    }
    CU(cudaEventRecord(h->ev[1], h->stream));
    h->parents_done = true;
    return GMS_OK;
}

int compute_crosses_impl(gms_handle* h, int sync) {
    if (!h->parents_done) return fail(h, GMS_ESTATE, "gms_compute_parents has not been called");
    if (int rc = bind(h)) return rc;
    const int ncls = h->ncomp, TT = h->T * h->T;
    if (h->par_parts > 1 && !h->todo.empty()) {
        double* tabs[3] = {h->cls[0].table.p, h->cls[1].table.p, h->cls[2].table.p};
        CU(launch_unpack_rows(h->xbuf.p, ncls, TT, h->todo_d.p, (long long)h->todo.size(), tabs, h->stream));
        h->st.last_launches += 1;
    }
    CU(cudaEventRecord(h->ev[6], h->stream));
    const EffClass& dm = h->cls[1];
    bool timed_kernel = false;
    if (h->model >= GMS_MODEL_AD && h->N > 0) {
        // per-cross dominance term X = (d o xi)' RoR (d o xi), xi = delta_sire o delta_dam
        if (h->eff_mode == GMS_MODE_MAP_SCAN) {
            if (int rc = run_scan(h, dm.S.p, MASK_XI, h->sire_d.p, h->dam_d.p, h->N, h->scan_split_x, h->split_scan_x_d.p, h->X.p,
                                  nullptr, h->ev[4], h->ev[5])) return rc;
            timed_kernel = true;
        } else if (h->use_i8) {
            // in passes of at most i8_chunk crosses (the xi mask image of a pass is chunk x mp bytes)
            const bool one_pass = h->N <= h->i8_chunk;          // then the events bracket the kernel alone
            if (!one_pass) CU(cudaEventRecord(h->ev[4], h->stream));
            // the kernel walks every pass in sire order (stage_impl); finalize_sym puts row r of the pass where cross perm[r] of
            // the pass belongs, so X - and everything after it - stays in list order
            const int* ua = h->i8_sort ? h->sire_s_d.p : h->sire_d.p;
            const int* ub = h->i8_sort ? h->dam_s_d.p : h->dam_d.p;
            int pass = 0;
            for (long long off = 0; off < h->N; off += h->i8_chunk, ++pass) {
                const long long n = std::min<long long>(h->i8_chunk, h->N - off);
                const bool tail = n < h->i8_chunk && h->N > h->i8_chunk;
                if (!(one_pass && h->xi_image_valid && h->cache_on)) {       // cache off: every compute rebuilds every operand image
                    CU(launch_i8_build_xi(h->planeA.p, h->planeB.p, h->wpp, ua + off, ub + off, n, MASK_XI,
                                          h->mp_total / 64, h->xi_d.p, h->stream));
                    h->st.last_launches += 1;
                }
                double* Xp = h->X.p + off * (long long)TT;
                if (int rc = run_i8(h, dm, h->xi_d.p, MASK_XI, ua + off, ub + off, n,
                                    tail ? h->plan_i8_tail : h->plan_i8, tail ? h->split_i8_tail_d.p : h->split_i8_d.p, Xp,
                                    h->i8_sort ? h->perm_d.p + off : nullptr,
                                    one_pass ? h->ev[4] : nullptr, one_pass ? h->ev[5] : nullptr)) return rc;
                if (int rc = verify_and_repair(h, dm, h->R2.p, MASK_XI, h->sire_d.p + off, h->dam_d.p + off, n, Xp, false, 3 + pass)) return rc;
            }
            h->xi_image_valid = one_pass;
            if (!one_pass) CU(cudaEventRecord(h->ev[5], h->stream));
            timed_kernel = true;
        } else {
            if (int rc = run_contract(h, h->R2.p, dm.emat.p, MASK_XI, h->sire_d.p, h->dam_d.p, h->N, h->plan_x, h->split_x_d.p,
                                      h->Z.p, h->X.p, nullptr, nullptr, h->ev[4], h->ev[5])) return rc;
            timed_kernel = true;
        }
    }
    if (!timed_kernel) { CU(cudaEventRecord(h->ev[4], h->stream)); CU(cudaEventRecord(h->ev[5], h->stream)); }
    CU(cudaEventRecord(h->ev[2], h->stream));
    if (h->N > 0) {
        const long long w0 = h->chrom[h->c_begin].pad_off / 32;
        const long long w1 = ((long long)h->chrom[h->c_end - 1].pad_off + h->chrom[h->c_end - 1].mp) / 32;
        CU(launch_nseg(h->planeA.p, h->planeB.p, h->wpp, w0, w1, h->sire_d.p, h->dam_d.p, h->N, h->nseg_d.p, h->stream));
        AssembleParams ap{};
        ap.QA = h->cls[0].table.p; ap.PD = h->cls[1].table.p; ap.QB = h->cls[2].table.p; ap.X = h->X.p;
        ap.sire = h->sire_d.p; ap.dam = h->dam_d.p;
        ap.pair_t1 = h->pair1_d.p; ap.pair_t2 = h->pair2_d.p;
        ap.N = h->N; ap.T = h->T; ap.npairs = h->npairs; ap.ncomp = h->ncomp; ap.out = h->out_d.p;
        CU(launch_assemble(ap, h->stream));
        h->st.last_launches += 2;
    }
    CU(cudaEventRecord(h->ev[3], h->stream));
    // the rows computed by this call (on any rank) are valid for the cached effects from now on
    for (int p : h->todo) h->have[p] = 1;
    h->cache_valid = h->cache_on;
    h->computed = true;
    h->settled = false;
    if (sync) CU(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

// After a synchronisation: how many units failed the int8 bound; if more than the in-stream lists hold, redo the whole
// call with the FP64 contraction (every rank recomputes all its parents then: no second exchange is needed).
int settle(gms_handle* h) {
    if (h->settled) return GMS_OK;
    h->settled = true;
    h->st.last_fallback_units = 0;
    h->st.last_fallback_overflow = 0;
    if (!h->use_i8 || h->n_counters == 0 || h->tau <= 0.0) return GMS_OK;
    h->counters_h.assign(h->n_counters, 0);
    CU(cudaMemcpy(h->counters_h.data(), h->fb_counter.p, sizeof(int) * h->n_counters, cudaMemcpyDeviceToHost));
    bool over = false;
    for (int c : h->counters_h) { h->st.last_fallback_units += c; over |= c > h->fb_cap; }
    if (!over) return GMS_OK;
    h->st.last_fallback_overflow = 1;
    h->overflow_key = h->raw_key;                   // GMS_MODE_AUTO goes straight to FP64 when these effects come again
    h->overflow_valid = true;
    h->eff_mode = GMS_MODE_DENSE;
    h->use_i8 = false;
    h->n_counters = 0;
    h->cache_valid = false;
    h->have.assign((size_t)h->P, 0);
    h->todo.clear();
    std::vector<unsigned char> need((size_t)h->P, 0);
    std::vector<int> s((size_t)h->N), d((size_t)h->N);
    // the parents this rank's crosses need (the exchange buffer is not reused)
    if (h->N) {
        CU(cudaMemcpy(s.data(), h->sire_d.p, sizeof(int) * h->N, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(d.data(), h->dam_d.p, sizeof(int) * h->N, cudaMemcpyDeviceToHost));
    }
    for (long long x = 0; x < h->N; ++x) need[s[x]] = need[d[x]] = 1;
    for (long long p = 0; p < h->P; ++p) if (need[p]) h->todo.push_back((int)p);
    const int saved_parts = h->par_parts, saved_index = h->par_index;
    h->par_parts = 1; h->par_index = 0;
    h->todo_lo = 0; h->todo_n = h->todo_L = (long long)h->todo.size();
    int rc = GMS_OK;
    do {
        const int TT = h->T * h->T, NU = BN / h->T;
        if (cudaSuccess != h->todo_d.ensure(h->todo.size() + 1)) { rc = fail(h, GMS_ENOMEM, "allocation failed"); break; }
        if (!h->todo.empty()) cudaMemcpy(h->todo_d.p, h->todo.data(), sizeof(int) * h->todo.size(), cudaMemcpyHostToDevice);
        make_plan(h, h->todo_n, NU, h->plan_par);
        make_plan(h, h->model >= GMS_MODEL_AD ? h->N : 0, NU, h->plan_x);
        if ((rc = upload_plan(h, h->plan_par, h->split_par_d))) break;
        if ((rc = upload_plan(h, h->plan_x, h->split_x_d))) break;
        const size_t zneed = std::max<size_t>((size_t)h->plan_par.nsplit * (size_t)h->todo_n * TT,
                                              h->model >= GMS_MODEL_AD ? (size_t)h->plan_x.nsplit * (size_t)h->N * TT : (size_t)0);
        if (cudaSuccess != h->Z.ensure(zneed)) { rc = fail(h, GMS_ENOMEM, "allocation failed"); break; }
        h->prep_needed = false;                     // the padded effect matrices are in place
        if ((rc = compute_parents_impl(h))) break;
        if ((rc = compute_crosses_impl(h, 1))) break;
    } while (0);
    h->par_parts = saved_parts; h->par_index = saved_index;
    h->cache_valid = false;                         // the int8 operand images no longer match the resolved mode
    h->settled = true;
    return rc;
}

}  // namespace

extern "C" {

const char* gms_version(void) {
    static const std::string v = std::string("gms_b200 0.3 (sm_100a; int8 tcgen05 contraction, digit planes: ") + std::to_string(I8_DIG) +
                                 ", FP64 error-bound fallback; FP64 DMMA contraction; map scan)";
    return v.c_str();
}

const char* gms_last_error(const gms_handle* h) {
    if (h) return h->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mu);
    static thread_local std::string copy;
    copy = g_create_err;
    return copy.c_str();
}

int gms_create(gms_handle** out, int device) {
    gms_handle* h = nullptr;
    if (!out) return fail(h, GMS_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(h, GMS_ECUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(h, GMS_EINVAL, "device index out of range");
    gms_handle* nh = new (std::nothrow) gms_handle();
    if (!nh) return fail(h, GMS_ENOMEM, "host allocation failed");
    nh->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete nh;
        return fail(h, GMS_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        delete nh;
        return fail(h, GMS_ECUDA, "device is not sm_100-class (this library is built for sm_100a only)");
    }
    nh->num_sms = prop.multiProcessorCount;
    nh->smem_limit = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&nh->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete nh;
        return fail(h, GMS_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    nh->stream = nh->own_stream;
    if (const char* ev = getenv("GMS_I8_CG2")) nh->i8_pair = atoi(ev) != 0;      // test hook: one CTA per 128 crosses
    if (const char* ev = getenv("GMS_I8_SORT")) nh->i8_sort = atoi(ev) != 0;     // test hook: list order
#ifdef GMS_PROFILE
    if (const char* ev = getenv("GMS_I8_DEBUG")) nh->i8_debug = atoi(ev);        // profiling builds only: results are garbage
#endif
    for (auto& ev : nh->ev) cudaEventCreate(&ev);
    *out = nh;
    return GMS_OK;
}

void gms_destroy(gms_handle* h) {
    if (!h) return;
    {
        LOCK(h);
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        for (auto& c : h->cls) c.release();
        h->chrom_d.release(); h->R.release(); h->R2.release(); h->planeA.release(); h->planeB.release();
        h->rowtiles_d.release(); h->sire_d.release(); h->dam_d.release(); h->sire_s_d.release(); h->dam_s_d.release(); h->perm_d.release(); h->pair1_d.release(); h->pair2_d.release();
        h->split_par_d.release(); h->split_x_d.release(); h->split_fb_d.release(); h->flag_d.release();
        h->Z.release(); h->Zfb.release(); h->X.release(); h->out_d.release(); h->gebv_d.release(); h->mean_d.release();
        h->cm_add.release(); h->cm_dom.release(); h->cm_raw.release(); h->tidy_d.release(); h->nseg_d.release(); h->todo_d.release(); h->xbuf.release();
        h->jtiles_d.release(); h->gbase_d.release(); h->xi_d.release(); h->xi_het_d.release(); h->xi_delta_d.release();
        h->split_i8_d.release(); h->split_i8_par_d.release(); h->split_i8_tail_d.release(); h->i8_blkpref_d.release();
        h->fb_counter.release(); h->fb_a.release(); h->fb_b.release(); h->fb_row.release(); h->ver_n.release(); h->ver_b.release();
        h->scan_ab.release(); h->split_scan_par_d.release(); h->split_scan_x_d.release();
        for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
        if (h->own_stream) cudaStreamDestroy(h->own_stream);
    }
    delete h;
}

int gms_set_stream(gms_handle* h, void* cuda_stream) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (int rc = bind(h)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return GMS_OK;
}

int gms_set_genmap(gms_handle* h, int C, const int64_t* m_c, const double* cM) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!cM) return fail(h, GMS_EINVAL, "cM is NULL");
    if (C > 0 && m_c) {
        long long m = 0;
        for (int c = 0; c < C; ++c) m += m_c[c] > 0 ? m_c[c] : 0;
        for (long long j = 0; j < m; ++j)
            if (!std::isfinite(cM[j])) return fail(h, GMS_EINVAL, "non-finite map position");
    }
    if (int rc = bind(h)) return rc;
    if (int rc = set_geometry(h, C, m_c)) return rc;
    DevBuf<double> cm_d;
    CU(cm_d.ensure((size_t)h->m));
    CU(cudaMemcpyAsync(cm_d.p, cM, sizeof(double) * h->m, cudaMemcpyHostToDevice, h->stream));
    for (int c = 0; c < C; ++c)
        CU(launch_build_tiles_from_map(cm_d.p + h->chrom[c].snp_off, h->chrom[c], h->R.p, h->R2.p, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cm_d.release();
    // map-structured fast path: a_j = exp(-c x_j), b_j = exp(c x_j) with x relative to the chromosome start;
    // usable when positions are sorted within every chromosome and c * length stays far from overflow
    h->scan_ok = true;
    std::vector<double> ab((size_t)4 * h->mp_total, 0.0);
    for (int c = 0; c < C && h->scan_ok; ++c) {
        const ChromDesc& cd = h->chrom[c];
        const double x0 = cM[cd.snp_off];
        for (int j = 0; j < cd.m; ++j) {
            const double x = cM[cd.snp_off + j];
            if (j > 0 && x < cM[cd.snp_off + j - 1]) { h->scan_ok = false; break; }
            const double d = x - x0;
            if (4.0 * (d / 100.0) > 600.0) { h->scan_ok = false; break; }
            const size_t pj = (size_t)cd.pad_off + j;
            ab[0 * h->mp_total + pj] = exp(-2.0 * (d / 100.0));
            ab[1 * h->mp_total + pj] = exp(2.0 * (d / 100.0));
            ab[2 * h->mp_total + pj] = exp(-4.0 * (d / 100.0));
            ab[3 * h->mp_total + pj] = exp(4.0 * (d / 100.0));
        }
    }
    if (h->scan_ok) {
        CU(h->scan_ab.ensure(ab.size()));
        CU(cudaMemcpy(h->scan_ab.p, ab.data(), ab.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (h->mode == GMS_MODE_MAP_SCAN) h->mode = GMS_MODE_AUTO;
    h->have_tiles = true;
    return build_worklist(h);
}

int gms_set_mode(gms_handle* h, int mode) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (mode != GMS_MODE_DENSE && mode != GMS_MODE_MAP_SCAN && mode != GMS_MODE_INT8_TENSOR && mode != GMS_MODE_AUTO)
        return fail(h, GMS_EINVAL, "unknown mode");
    if (mode == GMS_MODE_INT8_TENSOR && !h->have_tiles) return fail(h, GMS_ESTATE, "set the map / recombination blocks first");
    if (mode == GMS_MODE_MAP_SCAN && !(h->have_tiles && h->scan_ok))
        return fail(h, GMS_ESTATE, "the map-structured fast path needs gms_set_genmap() with positions sorted within every chromosome");
    h->mode = mode;
    invalidate(h);
    return GMS_OK;
}

int gms_set_int8_tolerance(gms_handle* h, double tau) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!(tau == tau) || tau > 1.0) return fail(h, GMS_EINVAL, "tolerance must be <= 1 (<= 0 disables the error bound)");
    h->tau = tau;
    invalidate(h);
    return GMS_OK;
}

int gms_set_cache(gms_handle* h, int enabled) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    h->cache_on = enabled != 0;
    h->cache_valid = false;
    h->xi_image_valid = false;
    return GMS_OK;
}

int gms_set_parent_shard(gms_handle* h, int index, int parts) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (parts < 1 || index < 0 || index >= parts) return fail(h, GMS_EINVAL, "bad parent shard");
    h->par_index = index;
    h->par_parts = parts;
    h->staged = h->parents_done = h->computed = false;
    return GMS_OK;
}

int gms_set_recomb_blocks(gms_handle* h, int C, const int64_t* m_c, const double* const* blocks) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!blocks) return fail(h, GMS_EINVAL, "blocks is NULL");
    long long mmax = 0;
    for (int c = 0; c < C && m_c; ++c) {
        if (!blocks[c]) return fail(h, GMS_EINVAL, "blocks[c] is NULL");
        mmax = std::max<long long>(mmax, m_c[c]);
    }
    if (int rc = bind(h)) return rc;
    if (int rc = set_geometry(h, C, m_c)) return rc;
    DevBuf<double> blk_d;
    CU(blk_d.ensure((size_t)mmax * mmax));
    cudaError_t e = h->flag_d.ensure(4);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->flag_d.p, 0, sizeof(int) * 4, h->stream);
    for (int c = 0; c < C && e == cudaSuccess; ++c) {
        const size_t n = (size_t)m_c[c] * m_c[c];
        e = cudaMemcpyAsync(blk_d.p, blocks[c], n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = launch_build_tiles_from_block(blk_d.p, h->chrom[c], h->R.p, h->R2.p, h->flag_d.p, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);   // blk_d is reused
    }
    int asym = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&asym, h->flag_d.p, sizeof(int), cudaMemcpyDeviceToHost);
    blk_d.release();
    CU(e);
    if (asym) {
        char buf[160];
        snprintf(buf, sizeof buf, "recombFreqMat block is not symmetric / finite (%d entries differ from their transpose)", asym);
        return fail(h, GMS_EINVAL, buf);
    }
    h->scan_ok = false;
    if (h->mode == GMS_MODE_MAP_SCAN) h->mode = GMS_MODE_AUTO;
    h->have_tiles = true;
    return build_worklist(h);
}

static int set_haps_common(gms_handle* h, long long P, long long m, const void* hap, bool i32cm) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!h->have_tiles) return fail(h, GMS_ESTATE, "set the map / recombination blocks before the haplotypes");
    if (!hap || P <= 0) return fail(h, GMS_EINVAL, "hap is NULL or P <= 0");
    if (m != h->m) return fail(h, GMS_EINVAL, "haplotype matrix has a different number of SNPs than the map");
    if (P > (1LL << 24)) return fail(h, GMS_EINVAL, "too many parents (at most 2^24)");
    if (int rc = bind(h)) return rc;
    invalidate(h);
    h->have_haps = false;
    CU(h->planeA.ensure((size_t)P * h->wpp));
    CU(h->planeB.ensure((size_t)P * h->wpp));
    CU(h->flag_d.ensure(4));
    CU(cudaMemsetAsync(h->flag_d.p, 0, sizeof(int) * 4, h->stream));
    const size_t elems = (size_t)2 * P * m;
    void* raw = nullptr;
    CU(cudaMalloc(&raw, elems * (i32cm ? 4 : 1)));
    cudaError_t e = cudaMemcpyAsync(raw, hap, elems * (i32cm ? 4 : 1), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess)
        e = i32cm ? launch_pack_i32cm((const int32_t*)raw, P, m, h->chrom_d.p, h->C, h->planeA.p, h->planeB.p, h->wpp,
                                      h->flag_d.p, h->stream)
                  : launch_pack_u8rm((const uint8_t*)raw, P, m, h->chrom_d.p, h->C, h->planeA.p, h->planeB.p, h->wpp,
                                     h->flag_d.p, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(raw);
    CU(e);
    int bad = 0;
    CU(cudaMemcpy(&bad, h->flag_d.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) {
        char buf[128];
        snprintf(buf, sizeof buf, "haploMat has %d values outside {0,1}", bad);
        return fail(h, GMS_EINVAL, buf);
    }
    h->P = P;
    h->have.assign((size_t)P, 0);
    h->have_haps = true;
    return GMS_OK;
}

int gms_set_haplotypes_i32cm(gms_handle* h, int64_t P, int64_t m, const int32_t* hap) {
    return set_haps_common(h, P, m, hap, true);
}
int gms_set_haplotypes_u8rm(gms_handle* h, int64_t P, int64_t m, const uint8_t* hap) {
    return set_haps_common(h, P, m, hap, false);
}

int gms_set_chromosome_shard(gms_handle* h, int c_begin, int c_end) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!h->have_tiles) return fail(h, GMS_ESTATE, "set the map / recombination blocks first");
    if (c_begin < 0 || c_end > h->C || c_begin >= c_end) return fail(h, GMS_EINVAL, "bad chromosome range");
    if (int rc = bind(h)) return rc;
    h->c_begin = c_begin;
    h->c_end = c_end;
    invalidate(h);
    return build_worklist(h);
}

int gms_trait_pairs(int T, int32_t* t1, int32_t* t2) {
    if (T < 1) return GMS_EINVAL;
    int n = 0;
    for (int t = 0; t < T; ++t, ++n) if (t1 && t2) { t1[n] = t; t2[n] = t; }
    for (int a = 0; a < T; ++a)
        for (int b = a + 1; b < T; ++b, ++n) if (t1 && t2) { t1[n] = a; t2[n] = b; }
    return n;
}

int gms_stage_inputs(gms_handle* h, int model, int T, const double* add, const double* dom, const double* asub,
                     int64_t N, const int32_t* sire, const int32_t* dam) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    return stage_impl(h, model, T, add, dom, asub, N, sire, dam, 0, N);
}

int gms_stage_inputs_range(gms_handle* h, int model, int T, const double* add, const double* dom, const double* asub,
                           int64_t N, const int32_t* sire, const int32_t* dam, int64_t x_begin, int64_t x_end) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    return stage_impl(h, model, T, add, dom, asub, N, sire, dam, x_begin, x_end);
}

int gms_compute_parents(gms_handle* h) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    return compute_parents_impl(h);
}

int gms_parent_exchange(gms_handle* h, void** dev_buf, int64_t* doubles_per_part, int32_t* parts) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!h->parents_done) return fail(h, GMS_ESTATE, "gms_compute_parents has not been called");
    const bool any = h->par_parts > 1 && !h->todo.empty();
    if (dev_buf) *dev_buf = any ? (void*)h->xbuf.p : nullptr;
    if (doubles_per_part) *doubles_per_part = any ? (int64_t)h->todo_L * h->ncomp * h->T * h->T : 0;
    if (parts) *parts = h->par_parts;
    return GMS_OK;
}

int gms_compute_crosses(gms_handle* h, int sync) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (int rc = compute_crosses_impl(h, sync)) return rc;
    return sync ? settle(h) : GMS_OK;
}

int gms_compute(gms_handle* h, int sync) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (h->par_parts > 1) return fail(h, GMS_ESTATE, "a parent-sharded handle needs gms_compute_parents / exchange / gms_compute_crosses");
    if (int rc = compute_parents_impl(h)) return rc;
    if (int rc = compute_crosses_impl(h, sync)) return rc;
    return sync ? settle(h) : GMS_OK;
}

int gms_fetch_outputs(gms_handle* h, double* out, int32_t* nseg) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!h->computed) return fail(h, GMS_ESTATE, "gms_compute has not been called");
    if (int rc = bind(h)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    if (int rc = settle(h)) return rc;
    if (h->N) {
        if (out) CU(cudaMemcpyAsync(out, h->out_d.p, sizeof(double) * h->N * h->npairs * h->ncomp, cudaMemcpyDeviceToHost, h->stream));
        if (nseg) CU(cudaMemcpyAsync(nseg, h->nseg_d.p, sizeof(int) * h->N, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

int gms_device_outputs(gms_handle* h, void** out_dev, void** nseg_dev, int64_t* n_rows) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!h->computed) return fail(h, GMS_ESTATE, "gms_compute has not been called");
    if (out_dev) *out_dev = h->N ? (void*)h->out_d.p : nullptr;
    if (nseg_dev) *nseg_dev = h->N ? (void*)h->nseg_d.p : nullptr;
    if (n_rows) *n_rows = h->N;
    return GMS_OK;
}

int gms_predict(gms_handle* h, int model, int T, const double* add, const double* dom, const double* asub,
                int64_t N, const int32_t* sire, const int32_t* dam, double* out, int32_t* nseg) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (N > 0 && (!out || !nseg)) return fail(h, GMS_EINVAL, "out / nseg is NULL");
    if (h->par_parts > 1) return fail(h, GMS_ESTATE, "gms_predict on a parent-sharded handle: use gms_predict_sharded or the split calls");
    if (int rc = stage_impl(h, model, T, add, dom, asub, N, sire, dam, 0, N)) return rc;
    if (int rc = compute_parents_impl(h)) return rc;
    if (int rc = compute_crosses_impl(h, 0)) return rc;
    return gms_fetch_outputs(h, out, nseg);
}

int gms_predict_batched(gms_handle* h, int model, int nsets, int T, const double* add, const double* dom, const double* asub,
                        int64_t N, const int32_t* sire, const int32_t* dam, double* out, int32_t* nseg) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (nsets < 1) return fail(h, GMS_EINVAL, "nsets must be >= 1");
    if (N > 0 && (!out || !nseg)) return fail(h, GMS_EINVAL, "out / nseg is NULL");
    if (model < GMS_MODEL_A || model > GMS_MODEL_DIRDOM || T < 1) return fail(h, GMS_EINVAL, "bad model / T");
    const size_t es = (size_t)h->m * T, os = (size_t)N * (T * (T + 1) / 2) * (model + 1);
    for (int s = 0; s < nsets; ++s) {
        // one staged cross list for all sets: the cross-side operands (sire / dam lists, the int8 path's xi image, Nsegsnps)
        // are keyed on the cross list and reused; only what depends on the effects is rebuilt per set
        if (int rc = gms_predict(h, model, T, add + s * es, dom ? dom + s * es : nullptr, asub ? asub + s * es : nullptr, N, sire, dam,
                                 out + s * os, nseg))
            return rc;
    }
    return GMS_OK;
}

int gms_predict_sharded(gms_handle* const* hs, int nh, int model, int T, const double* add, const double* dom,
                        const double* asub, int64_t N, const int32_t* sire, const int32_t* dam, double* out,
                        int32_t* nseg) {
    if (!hs || nh < 1 || !hs[0]) return GMS_EINVAL;
    for (int i = 0; i < nh; ++i) if (!hs[i]) return fail(hs[0], GMS_EINVAL, "NULL handle in gms_predict_sharded");
    if (model < GMS_MODEL_A || model > GMS_MODEL_DIRDOM || T < 1) return fail(hs[0], GMS_EINVAL, "bad model / T");
    if (N < 0 || (N > 0 && (!sire || !dam || !out || !nseg))) return fail(hs[0], GMS_EINVAL, "bad cross list / outputs");
    const int64_t stride = (int64_t)(T * (T + 1) / 2) * (model + 1);
    std::vector<int> rc((size_t)nh, GMS_OK);
    auto range = [&](int i, int64_t& lo, int64_t& n) {
        const int64_t base = N / nh, extra = N % nh;
        lo = i * base + std::min<int64_t>(i, extra);
        n = base + (i < extra ? 1 : 0);
    };
    auto run_all = [&](auto fn) {
        std::vector<std::thread> th;
        for (int i = 1; i < nh; ++i) th.emplace_back([&, i]() { if (rc[i] == GMS_OK) rc[i] = fn(i); });
        if (rc[0] == GMS_OK) rc[0] = fn(0);
        for (auto& t : th) t.join();
    };
    // phase 1: every handle stages the whole list (same parent set everywhere), computes ITS slice of the per-parent tables
    run_all([&](int i) -> int {
        gms_handle* h = hs[i];
        LOCK(h);
        const int sp = h->par_parts, si = h->par_index;
        h->par_parts = nh; h->par_index = i;
        int64_t lo, n;
        range(i, lo, n);
        int r = stage_impl(h, model, T, add, dom, asub, N, sire, dam, lo, lo + n);
        if (r == GMS_OK) r = compute_parents_impl(h);
        if (r == GMS_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) r = fail(h, GMS_ECUDA, "stream synchronise failed");
        if (r != GMS_OK) { h->par_parts = sp; h->par_index = si; }
        return r;
    });
    bool ok = true;
    for (int i = 0; i < nh; ++i) ok &= rc[i] == GMS_OK;
    // phase 2: all-gather of the table slices over peer copies, then the per-cross stage; results straight into the caller's slab
    if (ok)
        run_all([&](int i) -> int {
            gms_handle* h = hs[i];
            LOCK(h);
            int r = bind(h);
            if (r == GMS_OK && nh > 1 && !h->todo.empty()) {
                const size_t part = (size_t)h->todo_L * h->ncomp * h->T * h->T;
                for (int j = 0; j < nh && r == GMS_OK; ++j) {
                    if (j == i) continue;
                    if (cudaMemcpyPeerAsync(h->xbuf.p + j * part, h->device, hs[j]->xbuf.p + j * part, hs[j]->device,
                                            part * sizeof(double), h->stream) != cudaSuccess)
                        r = fail(h, GMS_ECUDA, "peer copy of the per-parent tables failed");
                }
            }
            if (r == GMS_OK) r = compute_crosses_impl(h, 0);
            int64_t lo, n;
            range(i, lo, n);
            if (r == GMS_OK) r = gms_fetch_outputs(h, out + lo * stride, nseg + lo);
            return r;
        });
    for (int i = 0; i < nh; ++i) { LOCK(hs[i]); hs[i]->par_parts = 1; hs[i]->par_index = 0; hs[i]->staged = hs[i]->parents_done = hs[i]->computed = false; }
    for (int i = 0; i < nh; ++i)
        if (rc[i] != GMS_OK) {
            if (i != 0) hs[0]->err = "shard " + std::to_string(i) + ": " + hs[i]->err;
            return rc[i];
        }
    return GMS_OK;
}

int gms_selind(int model, int T, int64_t N, const double* out, const double* w, double* si) {
    if (model < GMS_MODEL_A || model > GMS_MODEL_DIRDOM || T < 1 || N < 0 || !out || !w || !si) return GMS_EINVAL;
    const int ncomp = model + 1, npairs = T * (T + 1) / 2;
    std::vector<int32_t> p1(npairs), p2(npairs);
    gms_trait_pairs(T, p1.data(), p2.data());
    for (int64_t x = 0; x < N; ++x)
        for (int c = 0; c < ncomp; ++c) {
            double v = 0.0;
            for (int pr = 0; pr < npairs; ++pr) {
                const double g = out[(x * npairs + pr) * ncomp + c];
                v += (p1[pr] == p2[pr] ? 1.0 : 2.0) * w[p1[pr]] * w[p2[pr]] * g;   // G symmetric (R/gmsFunctions.R:841)
            }
            si[x * ncomp + c] = v;
        }
    return GMS_OK;
}

int gms_cross_means(gms_handle* h, int T, const double* add, const double* dom, int64_t N, const int32_t* sire,
                    const int32_t* dam, double* gebv, double* mean_bv, double* mean_tgv) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!h->have_tiles || !h->have_haps) return fail(h, GMS_ESTATE, "map and haplotypes must be set first");
    if (T < 1 || T > MAX_T || !add || N < 0 || (N > 0 && (!sire || !dam))) return fail(h, GMS_EINVAL, "bad arguments");
    if ((dom == nullptr) != (mean_tgv == nullptr)) return fail(h, GMS_EINVAL, "mean_tgv is required iff dom is given");
    for (int64_t x = 0; x < N; ++x)
        if (sire[x] < 0 || sire[x] >= h->P || dam[x] < 0 || dam[x] >= h->P) return fail(h, GMS_EINVAL, "parent index out of range");
    if (int rc = bind(h)) return rc;
    // own effect buffers: the staged problem (and the cached operand images) of the variance path stay untouched
    const size_t n = (size_t)h->m * T;
    CU(h->cm_raw.ensure(n));
    CU(h->cm_add.ensure((size_t)h->mp_total * T));
    CU(cudaMemcpyAsync(h->cm_raw.p, add, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(launch_scatter_effects(h->cm_raw.p, h->m, T, h->chrom_d.p, h->C, h->cm_add.p, h->mp_total, h->stream));
    if (dom) {
        CU(h->cm_dom.ensure((size_t)h->mp_total * T));
        CU(cudaMemcpyAsync(h->cm_raw.p, dom, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));   // stream order keeps the reuse safe
        CU(launch_scatter_effects(h->cm_raw.p, h->m, T, h->chrom_d.p, h->C, h->cm_dom.p, h->mp_total, h->stream));
    }
    CU(h->gebv_d.ensure((size_t)h->P * T));
    CU(launch_gebv(h->planeA.p, h->planeB.p, h->wpp, nullptr, h->P, h->cm_add.p, h->mp_total, T, h->gebv_d.p, h->stream));
    std::vector<double> g((size_t)h->P * T);
    CU(cudaMemcpyAsync(g.data(), h->gebv_d.p, sizeof(double) * g.size(), cudaMemcpyDeviceToHost, h->stream));
    if (dom && N) {
        DevBuf<int> s_d, d_d;
        cudaError_t e = s_d.ensure((size_t)N);
        if (e == cudaSuccess) e = d_d.ensure((size_t)N);
        if (e == cudaSuccess) e = h->mean_d.ensure((size_t)N * T);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s_d.p, sire, sizeof(int) * N, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_d.p, dam, sizeof(int) * N, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = launch_tgv_mean(h->planeA.p, h->planeB.p, h->wpp, s_d.p, d_d.p, N, h->cm_add.p, h->cm_dom.p,
                                                  h->mp_total, T, h->mean_d.p, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(mean_tgv, h->mean_d.p, sizeof(double) * N * T, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        s_d.release(); d_d.release();
        CU(e);
    }
    CU(cudaStreamSynchronize(h->stream));
    if (gebv) memcpy(gebv, g.data(), sizeof(double) * g.size());
    if (mean_bv)
        for (int64_t x = 0; x < N; ++x)
            for (int t = 0; t < T; ++t)
                mean_bv[x * T + t] = (g[(size_t)sire[x] * T + t] + g[(size_t)dam[x] * T + t]) / 2.0;   // R/predCrossVar.R:369
    return GMS_OK;
}

int gms_tidy(gms_handle* h, const double* si_weights, double std_sel_int, double* tidy) {
    if (!h) return GMS_EINVAL;
    LOCK(h);
    if (!h->computed) return fail(h, GMS_ESTATE, "gms_compute has not been called");
    if (!tidy && h->N > 0) return fail(h, GMS_EINVAL, "tidy is NULL");
    if (int rc = bind(h)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    if (int rc = settle(h)) return rc;
    if (h->N == 0) return GMS_OK;
    const int T = h->T, nof = h->model == GMS_MODEL_A ? 1 : 2, ntr = T + (si_weights ? 1 : 0);
    // means from the effects already in HBM: MeanBV from the allele-substitution class, MeanTGV (DirDom) from (add, dom)
    const EffClass& bvc = h->cls[h->model == GMS_MODEL_DIRDOM ? 2 : 0];
    CU(h->gebv_d.ensure((size_t)h->P * T));
    CU(launch_gebv(h->planeA.p, h->planeB.p, h->wpp, nullptr, h->P, bvc.emat.p, h->mp_total, T, h->gebv_d.p, h->stream));
    if (h->model == GMS_MODEL_DIRDOM) {
        CU(h->mean_d.ensure((size_t)h->N * T));
        CU(launch_tgv_mean(h->planeA.p, h->planeB.p, h->wpp, h->sire_d.p, h->dam_d.p, h->N, h->cls[0].emat.p, h->cls[1].emat.p,
                           h->mp_total, T, h->mean_d.p, h->stream));
    }
    CU(h->cm_raw.ensure((size_t)std::max<long long>(h->m * T, T)));     // scratch: the weights
    if (si_weights) CU(cudaMemcpyAsync(h->cm_raw.p, si_weights, sizeof(double) * T, cudaMemcpyHostToDevice, h->stream));
    const size_t n = (size_t)h->N * nof * ntr * 4;
    CU(h->tidy_d.ensure(n));
    TidyParams tp{};
    tp.out = h->out_d.p; tp.gebv = h->gebv_d.p; tp.mean_tgv = h->model == GMS_MODEL_DIRDOM ? h->mean_d.p : nullptr;
    tp.sire = h->sire_d.p; tp.dam = h->dam_d.p; tp.pair_t1 = h->pair1_d.p; tp.pair_t2 = h->pair2_d.p;
    tp.w = si_weights ? h->cm_raw.p : nullptr; tp.std_sel_int = std_sel_int;
    tp.N = h->N; tp.T = T; tp.npairs = h->npairs; tp.ncomp = h->ncomp; tp.model = h->model; tp.tidy = h->tidy_d.p;
    CU(launch_tidy(tp, h->stream));
    CU(cudaMemcpyAsync(tidy, h->tidy_d.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GMS_OK;
}

int gms_get_stats(gms_handle* h, gms_stats* s) {
    if (!h || !s) return GMS_EINVAL;
    LOCK(h);
    if (int rc = bind(h)) return rc;
    gms_stats& st = h->st;
    st.m = h->m; st.m_padded = h->mp_total; st.S2 = h->S2; st.tile_bytes = 2 * h->tile_doubles * 8;
    st.C = h->C; st.P = (int)h->P; st.last_T = h->T; st.last_model = h->model; st.last_N = h->N;
    st.last_parents_used = h->nP;
    st.last_parents_computed = h->todo_n;
    st.last_cache_hit = h->cache_hit ? 1 : 0;
    if (h->computed) {
        CU(cudaEventSynchronize(h->ev[3]));
        if (int rc = settle(h)) return rc;
        cudaEventElapsedTime(&st.last_ms_parent_stage, h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&st.last_ms_cross_stage, h->ev[6], h->ev[2]);
        cudaEventElapsedTime(&st.last_ms_assemble, h->ev[2], h->ev[3]);
        cudaEventElapsedTime(&st.last_ms_total, h->ev[0], h->ev[3]);
        st.last_ms_cross_kernel = 0.f;
        if (h->model >= GMS_MODEL_AD && h->N > 0) cudaEventElapsedTime(&st.last_ms_cross_kernel, h->ev[4], h->ev[5]);
        const double per_col = 2.0 * h->lower_elems;   // FMA = 2 flops, per (unit, trait) column
        st.last_flops_cross = h->model >= GMS_MODEL_AD ? per_col * (double)h->N * h->T : 0.0;
        st.last_mode = h->eff_mode;
        st.last_flops_parent = per_col * (double)h->todo_n * h->T * (h->model + 1);
    }
    *s = st;
    return GMS_OK;
}

}  // extern "C"
