// FP64 throughput probe for B200 (sm_100a): DFMA vs mma.sync f64 shapes.
// Decides which instruction the contraction kernel's mainloop is built on and gives
// the "what the pipe can do" ceiling next to the cuBLAS DGEMM ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double s) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    double a = s, b = 1.0 - s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int ILP>
__global__ void k_m8n8k4(double* out, int iters, double s) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = s, b = 1.0 - s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int ILP>
__global__ void k_m16n8k4(double* out, int iters, double s) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a0 = s, a1 = s * 0.5, b = 1.0 - s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int ILP>
__global__ void k_m16n8k8(double* out, int iters, double s) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a0 = s, a1 = s * 0.5, a2 = s * 0.25, a3 = s * 0.125, b0 = 1.0 - s, b1 = 0.5 - s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int ILP>
__global__ void k_m16n8k16(double* out, int iters, double s) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a0 = s, a1 = s * 0.5, a2 = s * 0.25, a3 = s * 0.125, a4 = s * 0.3, a5 = s * 0.7, a6 = s * 0.9, a7 = s * 0.1;
    double b0 = 1.0 - s, b1 = 0.5 - s, b2 = 0.25 - s, b3 = 0.125 - s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7),
                           "d"(b0), "d"(b1), "d"(b2), "d"(b3));
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
static int timeit(const char* name, F launch, double flops_per_thread_iter_x32, int threads, int blocks, int iters) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(iters / 8);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        launch(iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double warps = double(threads / 32) * blocks;
    double flops = warps * flops_per_thread_iter_x32 * iters;
    printf("%-26s threads=%4d blocks=%4d  %8.3f ms  %8.2f TFLOP/s\n", name, threads, blocks, best, flops / best * 1e-9);
    return 0;
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 4));
    const int iters = 20000;
    for (int threads : {128, 256, 512}) {
        int blocks = sms * (1024 / threads);
        constexpr int ILP = 8;
        // flops per WARP per iteration
        if (timeit("dfma ilp8", [&](int it) { k_dfma<ILP><<<blocks, threads>>>(out, it, 0.999); }, 32.0 * 2 * ILP, threads, blocks, iters)) return 1;
        if (timeit("mma.m8n8k4 ilp8", [&](int it) { k_m8n8k4<ILP><<<blocks, threads>>>(out, it, 0.999); }, 2.0 * 8 * 8 * 4 * ILP, threads, blocks, iters)) return 1;
        if (timeit("mma.m16n8k4 ilp8", [&](int it) { k_m16n8k4<ILP><<<blocks, threads>>>(out, it, 0.999); }, 2.0 * 16 * 8 * 4 * ILP, threads, blocks, iters)) return 1;
        if (timeit("mma.m16n8k8 ilp8", [&](int it) { k_m16n8k8<ILP><<<blocks, threads>>>(out, it, 0.999); }, 2.0 * 16 * 8 * 8 * ILP, threads, blocks, iters / 2)) return 1;
        if (timeit("mma.m16n8k16 ilp8", [&](int it) { k_m16n8k16<ILP><<<blocks, threads>>>(out, it, 0.999); }, 2.0 * 16 * 8 * 16 * ILP, threads, blocks, iters / 4)) return 1;
    }
    // low-occupancy points: what the contraction kernel actually runs at (8 math warps / SM)
    for (int threads : {256}) {
        int blocks = sms;
        if (timeit("mma.m16n8k4 ilp10 8w/SM", [&](int it) { k_m16n8k4<10><<<blocks, threads>>>(out, it, 0.999); }, 2.0 * 16 * 8 * 4 * 10, threads, blocks, iters)) return 1;
        if (timeit("mma.m16n8k8 ilp10 8w/SM", [&](int it) { k_m16n8k8<10><<<blocks, threads>>>(out, it, 0.999); }, 2.0 * 16 * 8 * 8 * 10, threads, blocks, iters / 2)) return 1;
        if (timeit("mma.m16n8k16 ilp10 8w/SM", [&](int it) { k_m16n8k16<10><<<blocks, threads>>>(out, it, 0.999); }, 2.0 * 16 * 8 * 16 * 10, threads, blocks, iters / 4)) return 1;
        if (timeit("dfma ilp16 8w/SM", [&](int it) { k_dfma<16><<<blocks, threads>>>(out, it, 0.999); }, 32.0 * 2 * 16, threads, blocks, iters)) return 1;
    }
    CK(cudaFree(out));
    return 0;
}
