// Probe 2: the contraction kernel's inner loop in isolation (fragment loads from shared memory +
// mma.m16n8k4.f64), no barriers, no producers: how close can W warps/SM with an MI x NI warp tile get to
// the DMMA pipe peak, with and without register double-buffering of the fragments?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop_probe dmma_loop_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b0));
}

template <int MI, int NI> struct Frag { double2 a[MI]; double b[NI]; };

template <int MI, int NI>
__device__ __forceinline__ void loadf(Frag<MI, NI>& f, const double* a, const double* b, int q) {
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) f.a[mi] = *reinterpret_cast<const double2*>(a + (q * 8 + 4 * mi) * 64);
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) f.b[ni] = b[(q * 10 + ni) * 32];
}
template <int MI, int NI>
__device__ __forceinline__ void mmaf(double (&acc)[MI][NI][4], const Frag<MI, NI>& f) {
#pragma unroll
    for (int ni = 0; ni < NI; ++ni)
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) dmma(acc[mi][ni], f.a[mi].x, f.a[mi].y, f.b[ni]);
}

template <int MI, int NI, bool PIPE>
__global__ void loopk(double* out, int iters) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;            // 4096 doubles
    double* Bs = sm + 4096;     // 2560 doubles
    for (int i = threadIdx.x; i < 4096 + 2560; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* a = As + ((warp & 3) * 32 + lane) * 2;
    const double* b = Bs + ((warp >> 2) % 2) * 5 * 32 + lane;
    double acc[MI][NI][4];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mi][ni][r] = 0.0;
    Frag<MI, NI> cur, nxt;
    if (PIPE) loadf<MI, NI>(cur, a, b, 0);
    for (int it = 0; it < iters; ++it) {
        if (PIPE) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                loadf<MI, NI>(nxt, a, b, (q + 1) & 7);
                mmaf<MI, NI>(acc, cur);
                cur = nxt;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                loadf<MI, NI>(cur, a, b, q);
                mmaf<MI, NI>(acc, cur);
            }
        }
    }
    double r = 0;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) r += acc[mi][ni][0] + acc[mi][ni][1] + acc[mi][ni][2] + acc[mi][ni][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MI, int NI, bool PIPE>
int run(const char* name, int warps, int sms, double* out) {
    const int iters = 4000;
    const size_t smem = (4096 + 2560) * 8;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(loopk<MI, NI, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    loopk<MI, NI, PIPE><<<sms, warps * 32, smem>>>(out, iters / 10);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        loopk<MI, NI, PIPE><<<sms, warps * 32, smem>>>(out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double flops = 2.0 * 16 * 8 * 4 * MI * NI * 8.0 * iters * warps * sms;
    printf("%-34s warps/SM=%2d  %8.3f ms  %7.2f TFLOP/s\n", name, warps, best, flops / best * 1e-9);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024));
    for (int w : {4, 8, 12, 16}) {
        if (run<2, 5, false>("tile 32x40 no-prefetch", w, sms, out)) return 1;
        if (run<2, 5, true>("tile 32x40 prefetch", w, sms, out)) return 1;
    }
    for (int w : {8, 16, 24}) {
        if (run<1, 5, true>("tile 16x40 prefetch", w, sms, out)) return 1;
        if (run<2, 3, true>("tile 32x24 prefetch", w, sms, out)) return 1;
    }
    for (int w : {4, 8}) if (run<4, 5, true>("tile 64x40 prefetch", w, sms, out)) return 1;
    CK(cudaFree(out));
    return 0;
}
