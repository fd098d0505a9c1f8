// Does CUDA-core arithmetic run concurrently with tcgen05 MMAs on B200?  One CTA per SM: warp 0 issues back-to-back
// tcgen05.mma.kind::i8 (M = 128, N = 224, K = 32, operands in shared memory, accumulator in TMEM), warps 1.. run a loop of
// independent FMA chains in FP64, FP32 or INT32.  Each part is timed alone and together (CUDA events).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_contention_probe tc_contention_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned long long desc(unsigned addr, unsigned lbo, unsigned sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3FFFu);
    d |= (unsigned long long)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

template <int KIND>   // 0 = FP64, 1 = FP32, 2 = INT32, 3 = one DFMA per 8 IMADs, 4 = one FFMA per 8 IMADs
__global__ void __launch_bounds__(32 * 17, 1) probe(int mma_batches, int alu_iters, int alu_warps, int skip_smsp0, double* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bar[2];
    __shared__ unsigned tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 * 32 + 224 * 32) / 4; i += blockDim.x)
        reinterpret_cast<unsigned*>(smem)[i] = 0x01ff0203u * (i * 2654435761u >> 7);       // arbitrary int8 pattern
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    if (warp == 0) {
        if (lane == 0) {
            const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24) | ((224u >> 3) << 17);
            const unsigned long long da = desc(smem_u32(smem), 128 * 16, 128), db = desc(smem_u32(smem) + 128 * 32, 28 * 128, 128);
            auto wait = [&](int b) {      // batch b committed to bar[b & 1]; its (b >> 1)-th use
                unsigned ok;
                do {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                                 : "=r"(ok) : "r"(smem_u32(&bar[b & 1])), "r"((unsigned)(b >> 1) & 1u) : "memory");
                } while (!ok);
            };
            for (int b = 0; b < mma_batches; ++b) {
                for (int i = 0; i < 16; ++i)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[b & 1])) : "memory");
                if (b >= 1) wait(b - 1);      // one batch in flight behind the one being issued
            }
            if (mma_batches > 0) wait(mma_batches - 1);
        }
    } else if (warp <= alu_warps && !(skip_smsp0 && (warp & 3) == 0)) {
        if (KIND == 0) {
            double a[8], x = 1.0 + 1e-9 * lane, y = 1e-12 * warp;
            for (int i = 0; i < 8; ++i) a[i] = i;
            for (int it = 0; it < alu_iters; ++it)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
            double s = 0;
            for (int i = 0; i < 8; ++i) s += a[i];
            if (s == 123.456) sink[threadIdx.x] = s;
        } else if (KIND == 1) {
            float a[8], x = 1.0f + 1e-6f * lane, y = 1e-7f * warp;
            for (int i = 0; i < 8; ++i) a[i] = i;
            for (int it = 0; it < alu_iters; ++it)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
            float s = 0;
            for (int i = 0; i < 8; ++i) s += a[i];
            if (s == 123.456f) sink[threadIdx.x] = s;
        } else if (KIND == 2) {
            int a[8], x = 3 + lane, y = 7 + warp;
            for (int i = 0; i < 8; ++i) a[i] = i;
            for (int it = 0; it < alu_iters; ++it)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = a[i] * x + y;
            int s = 0;
            for (int i = 0; i < 8; ++i) s += a[i];
            if (s == 123456) sink[threadIdx.x] = s;
        } else {
            int a[8], x = 3 + lane, y = 7 + warp;
            double d = 1.0, dx = 1.0 + 1e-9 * lane, dy = 1e-12 * warp;
            float f = 1.0f, fx = 1.0f + 1e-6f * lane, fy = 1e-7f * warp;
            for (int i = 0; i < 8; ++i) a[i] = i;
            for (int it = 0; it < alu_iters; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = a[i] * x + y;
                if (KIND == 3) d = fma(d, dx, dy); else f = fmaf(f, fx, fy);
            }
            int s = 0;
            for (int i = 0; i < 8; ++i) s += a[i];
            if (s == 123456 || d == 123.456 || f == 123.456f) sink[threadIdx.x] = s;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

template <int KIND>
static float run(int mma_batches, int alu_iters, int alu_warps, int skip, double* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = 128 * 32 + 224 * 32;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        probe<KIND><<<148, 32 * 17, smem>>>(mma_batches, alu_iters, alu_warps, skip, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); exit(1); }
    return ms;
}

int main() {
    double* sink;
    cudaMalloc(&sink, 4096 * sizeof(double));
    const int B = 6000;                 // 6000 batches x 16 MMAs x 112 cycles = 10.75 M tensor cycles (~5.5 ms at 1.965 GHz)
    const char* names[5] = {"FP64 DFMA", "FP32 FFMA", "INT32 IMAD", "8 IMAD + 1 DFMA", "8 IMAD + 1 FFMA"};
    const int iters[5] = {40000, 320000, 160000, 160000, 160000};
    printf("MMA alone (M128 N224 K32 kind::i8, %d x 16): %.3f ms  (ideal %.3f ms at 1.965 GHz)\n", B, run<0>(B, 0, 0, 0, sink),
           B * 16 * 112 / 1.965e6);
    for (int skip = 0; skip < 2; ++skip)
        for (int k = 0; k < 5; ++k)
            for (int w = 8; w <= 16; w += 8) {
                float alu = 0, both = 0;
                if (k == 0) { alu = run<0>(0, iters[k], w, skip, sink); both = run<0>(B, iters[k], w, skip, sink); }
                if (k == 1) { alu = run<1>(0, iters[k], w, skip, sink); both = run<1>(B, iters[k], w, skip, sink); }
                if (k == 2) { alu = run<2>(0, iters[k], w, skip, sink); both = run<2>(B, iters[k], w, skip, sink); }
                if (k == 3) { alu = run<3>(0, iters[k], w, skip, sink); both = run<3>(B, iters[k], w, skip, sink); }
                if (k == 4) { alu = run<4>(0, iters[k], w, skip, sink); both = run<4>(B, iters[k], w, skip, sink); }
                printf("%-16s %2d warps%s x %d: alone %.3f ms; together with the MMAs %.3f ms\n", names[k], w,
                       skip ? " (none on the MMA warp's scheduler)" : "", iters[k], alu, both);
            }
    return 0;
}
