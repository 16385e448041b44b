// Latency / throughput of the instructions the Crank-Nicolson scan is made of (sm_100a): dependent DFMA chain,
// independent DFMA streams, 32-bit SHFL (dependent and independent), __syncthreads, shared-memory round trip.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, long long *cyc, int n)
{
    double a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x * 1e-3 + j;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void k_shfl(int *out, long long *cyc, int n)
{
    int a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = __shfl_up_sync(0xffffffffu, a[j], 1) + 1;
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_bar(long long *cyc, int n)
{
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_smem(double *out, long long *cyc, int n)
{
    __shared__ double sm[1024];
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        sm[threadIdx.x] = v;
        __syncwarp();
        v = sm[threadIdx.x ^ 1] + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, 8);
    const int n = 4096;
#define RUN(name, kern, blocks, threads, per_iter)                                                              \
    kern<<<blocks, threads>>>(name##_args);                                                                     \
    cudaDeviceSynchronize();                                                                                    \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    int tcs[] = {32, 128, 256, 512, 1024};
    for (int tc : tcs) {
        k_dfma<1><<<148, tc>>>(out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dfma ilp1 threads/SM %4d: %.2f cyc/iter\n", tc, (double)h / n);
        k_dfma<2><<<148, tc>>>(out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dfma ilp2 threads/SM %4d: %.2f cyc/iter (%.2f per dfma)\n", tc, (double)h / n, (double)h / n / 2);
        k_dfma<4><<<148, tc>>>(out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dfma ilp4 threads/SM %4d: %.2f cyc/iter (%.2f per dfma)\n", tc, (double)h / n, (double)h / n / 4);
        k_dfma<8><<<148, tc>>>(out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dfma ilp8 threads/SM %4d: %.2f cyc/iter (%.2f per dfma)\n", tc, (double)h / n, (double)h / n / 8);
        k_shfl<1><<<148, tc>>>((int *)out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("shfl ilp1 threads/SM %4d: %.2f cyc/iter\n", tc, (double)h / n);
        k_shfl<4><<<148, tc>>>((int *)out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("shfl ilp4 threads/SM %4d: %.2f cyc/iter (%.2f per shfl)\n", tc, (double)h / n, (double)h / n / 4);
        k_shfl<8><<<148, tc>>>((int *)out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("shfl ilp8 threads/SM %4d: %.2f cyc/iter (%.2f per shfl)\n", tc, (double)h / n, (double)h / n / 8);
        k_bar<<<148, tc>>>(cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("bar       threads/SM %4d: %.2f cyc/iter\n", tc, (double)h / n);
        k_smem<<<148, tc>>>(out, cyc, n); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("smem rt   threads/SM %4d: %.2f cyc/iter\n", tc, (double)h / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
