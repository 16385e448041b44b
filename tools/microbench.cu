// Microbenchmarks that bound what a small (16 MB) in-place streaming pass can cost on this GPU:
// launch cadence, L2-resident read+write bandwidth at several grid shapes, FP64 sincos throughput.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void k_empty() {}
__global__ void k_scale(double2 *p, size_t n, double s)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double2 v = p[i];
        v.x *= s;
        v.y *= s;
        p[i] = v;
    }
}
// each thread handles M elements strided by T inside a "channel" of M*T elements (the engine's access pattern)
template <int M>
__global__ void k_scale_unit(double2 *p, int T, double s)
{
    double2 *base = p + (size_t)blockIdx.x * M * T * 2;  // pair of channels
    double2 a[M], b[M];
#pragma unroll
    for (int k = 0; k < M; ++k) a[k] = base[k * T + threadIdx.x];
#pragma unroll
    for (int k = 0; k < M; ++k) b[k] = base[M * T + k * T + threadIdx.x];
#pragma unroll
    for (int k = 0; k < M; ++k) {
        a[k].x = a[k].x * s + b[k].y;
        b[k].x = b[k].x * s - a[k].y;
    }
#pragma unroll
    for (int k = 0; k < M; ++k) base[k * T + threadIdx.x] = a[k];
#pragma unroll
    for (int k = 0; k < M; ++k) base[M * T + k * T + threadIdx.x] = b[k];
}
__global__ void k_sincos(double *out, int n_per_thread, double x0)
{
    double acc = 0, x = x0 + threadIdx.x * 1e-3 + blockIdx.x;
    for (int i = 0; i < n_per_thread; ++i) {
        double s, c;
        sincos(x, &s, &c);
        acc += s * c;
        x += 0.37;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_dfma(double *out, int n_per_thread)
{
    double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < n_per_thread; ++i) {
        a = fma(a, b, c);
        d = fma(d, b, c);
        e = fma(e, b, c);
        f = fma(f, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f;
}

template <typename F>
float time_it(F f, int reps)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / reps;  // us per call
}

int main()
{
    const size_t n = 1 << 20;  // 1M complex128 = 16 MB
    double2 *p;
    cudaMalloc(&p, n * sizeof(double2) * 16);
    cudaMemset(p, 0, n * sizeof(double2) * 16);
    double *out;
    cudaMalloc(&out, 148 * 2048 * sizeof(double));
    printf("empty kernel cadence: %.2f us\n", time_it([&] { k_empty<<<1, 32>>>(); }, 2000));
    printf("empty kernel 148x1024: %.2f us\n", time_it([&] { k_empty<<<148, 1024>>>(); }, 2000));
    for (int blocks : {148, 296, 592, 1184, 2368, 4096}) {
        for (int threads : {256, 512}) {
            float us = time_it([&] { k_scale<<<blocks, threads>>>(p, n, 1.0000001); }, 500);
            printf("scale 16MB grid %5d x %4d: %7.2f us  %7.1f GB/s (r+w)\n", blocks, threads, us, 2.0 * n * 16 / us * 1e-3);
        }
    }
    {
        float us = time_it([&] { k_scale<<<2368, 256>>>(p, n * 16, 1.0000001); }, 50);
        printf("scale 256MB (HBM): %7.2f us  %7.1f GB/s (r+w)\n", us, 2.0 * n * 16 * 16 / us * 1e-3);
    }
    {
        int T = 512;
        float us = time_it([&] { k_scale_unit<4><<<256, 512>>>(p, T, 1.0000001); }, 500);
        printf("unit pattern M=4 T=512 256 CTAs: %7.2f us  %7.1f GB/s\n", us, 2.0 * n * 16 / us * 1e-3);
        us = time_it([&] { k_scale_unit<4><<<512, 256>>>(p, 256, 1.0000001); }, 500);
        printf("unit pattern M=4 T=256 512 CTAs: %7.2f us  %7.1f GB/s\n", us, 2.0 * n * 16 / us * 1e-3);
        us = time_it([&] { k_scale_unit<2><<<512, 512>>>(p, 512, 1.0000001); }, 500);
        printf("unit pattern M=2 T=512 512 CTAs: %7.2f us  %7.1f GB/s\n", us, 2.0 * n * 16 / us * 1e-3);
        us = time_it([&] { k_scale_unit<2><<<2048, 128>>>(p, 128, 1.0000001); }, 500);
        printf("unit pattern M=2 T=128 2048 CTAs: %7.2f us  %7.1f GB/s\n", us, 2.0 * n * 16 / us * 1e-3);
        us = time_it([&] { k_scale_unit<8><<<256, 256>>>(p, 256, 1.0000001); }, 500);
        printf("unit pattern M=8 T=256 256 CTAs: %7.2f us  %7.1f GB/s\n", us, 2.0 * n * 16 / us * 1e-3);
    }
    {
        int npt = 64;
        float us = time_it([&] { k_sincos<<<148 * 4, 512>>>(out, npt, 0.3); }, 50);
        double rate = 148.0 * 4 * 512 * npt / us * 1e-3;  // G sincos/s
        printf("sincos f64: %.2f us -> %.2f G sincos/s\n", us, rate);
        us = time_it([&] { k_dfma<<<148 * 4, 512>>>(out, 4096); }, 20);
        printf("dfma: %.2f us -> %.2f TFLOP/s\n", us, 148.0 * 4 * 512 * 4096 * 4 * 2 / us * 1e-6);
    }
    // graph of 100 back-to-back small kernels
    {
        cudaStream_t st;
        cudaStreamCreate(&st);
        cudaGraph_t g;
        cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < 100; ++i) k_scale<<<1184, 256, 0, st>>>(p, n, 1.0000001);
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaGraphLaunch(ge, st);
        cudaEventRecord(e0, st);
        for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("graph: scale 16MB 1184x256 per kernel: %.2f us\n", ms * 1e3 / 1000);
        cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < 100; ++i) k_empty<<<148, 512, 0, st>>>();
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        cudaGraphLaunch(ge, st);
        cudaEventRecord(e0, st);
        for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("graph: empty 148x512 per kernel: %.2f us\n", ms * 1e3 / 1000);
    }
    return 0;
}
