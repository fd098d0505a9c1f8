// Inline PTX used by the int8 tcgen05 kernel (i8path.cu): mbarriers, bulk copies, UMMA descriptors, MMA issue /
// commit for cta_group::1 and ::2, cluster helpers.
#pragma once
#include <cuda_runtime.h>

namespace gms {

__device__ __forceinline__ unsigned i8_smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void i8_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void i8_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void i8_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void i8_mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void i8_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, no swizzle: uint128 offset of (row r0 + 8 r1, chunk c) = r0 + r1*SBO + c*LBO   (cute mma_traits_sm100.hpp)
__device__ __forceinline__ unsigned long long i8_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3FFFu);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                                  // descriptor version (Blackwell)
    return d;                                         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ void i8_mma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                       unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster. Relaxed: what the arrive publishes
// (bulk-copied operands already completed on the local barrier; TMEM reads fenced by tcgen05.fence) is already performed,
// and a release at cluster scope costs a full cluster fence per stage
__device__ __forceinline__ void i8_mbar_arrive_remote(unsigned bar, unsigned rank) {
    asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n}\n"
                 ::"r"(bar), "r"(rank) : "memory");
}
// one M = 256 MMA over a CTA pair: each CTA supplies its 128 A rows and HALF of the B rows from its own shared memory (same
// offsets in both CTAs); rows 0-127 of D land in the leader's TMEM, rows 128-255 in the peer's. Issued by the leader only.
__device__ __forceinline__ void i8_mma_cg2(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                           unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void i8_commit_cg2(unsigned bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void i8_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned i8_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void i8_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// one lane of a converged warp (the same lane every time for the full mask); the loop around it stays warp-uniform, so
// addresses, descriptors and counters live in uniform registers and each issue costs a handful of instructions
__device__ __forceinline__ bool v2_elect() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}


}  // namespace gms
