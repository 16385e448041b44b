// Neighbour-exchange microbenchmark for the on-chip resident kernel: 125 co-resident CTAs of 512 threads, each sends two
// "boundary channels" (4 complex128 per thread each) to its two neighbours and receives two, N rounds, no arithmetic.
// Variants:  0 LL (8-byte {payload32, seq} words) with .volatile accesses        (2x payload)
//            1 LL with .relaxed.gpu accesses
//            2 plain 16-byte stores + per-warp  fence + st.release flag / ld.acquire poll
//            3 as 2, one flag per CTA direction (bar.sync, one thread fences)
//            4 LL, one channel only per direction pair (half the volume: what an 8-channel block would exchange)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/xchg_bench tools/xchg_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define DEVINL __device__ __forceinline__

template <int SCOPE>
DEVINL void st4(uint4 *p, unsigned a, unsigned b, unsigned c, unsigned d)
{
    if (SCOPE == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    else asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int SCOPE>
DEVINL uint4 ld4(const uint4 *p)
{
    uint4 v;
    if (SCOPE == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
DEVINL void st_release(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
DEVINL unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// mailbox of CTA blk, direction dir (0 up, 1 down), slot: [8 units][T] uint4 (LL) or [4][T] double2 (plain)
template <int VAR>
__global__ void __launch_bounds__(512, 1) k_xchg(uint4 *halo, unsigned *flags, int rounds, int work, double *sink)
{
    const int T = blockDim.x, tid = threadIdx.x, kb = blockIdx.x, nblk = gridDim.x, lane = tid & 31, warp = tid >> 5;
    const bool has_lo = kb > 0, has_hi = kb + 1 < nblk;
    double2 A[4], D[4], GL[4], GR[4];
    for (int k = 0; k < 4; ++k) A[k] = make_double2(kb + k, tid), D[k] = make_double2(-kb - k, tid + 0.5);
    double acc = 0.0;
    const int NCH = (VAR == 4) ? 1 : 1;
    (void)NCH;
    for (int n = 1; n <= rounds; ++n) {
        const unsigned seq = (unsigned)n;
        if (VAR == 0 || VAR == 1 || VAR == 4) {
            constexpr int SC = (VAR == 1) ? 1 : 0;
            auto box = [&](int blk, int dir) { return halo + (((size_t)blk * 2 + dir) * 2 + (seq & 1u)) * 8 * T + tid; };
            const int nk = (VAR == 4) ? 2 : 4;
            if (has_lo)
                for (int k = 0; k < nk; ++k) {
                    st4<SC>(box(kb, 1) + (2 * k) * T, __double2loint(A[k].x), seq, __double2hiint(A[k].x), seq);
                    st4<SC>(box(kb, 1) + (2 * k + 1) * T, __double2loint(A[k].y), seq, __double2hiint(A[k].y), seq);
                }
            if (has_hi)
                for (int k = 0; k < nk; ++k) {
                    st4<SC>(box(kb, 0) + (2 * k) * T, __double2loint(D[k].x), seq, __double2hiint(D[k].x), seq);
                    st4<SC>(box(kb, 0) + (2 * k + 1) * T, __double2loint(D[k].y), seq, __double2hiint(D[k].y), seq);
                }
            // "interior work" between send and receive
            for (int i = 0; i < work; ++i) acc = fma(acc, 1.0000001, 1e-9);
            auto recv = [&](const uint4 *bx, double2(&v)[4]) {
                uint4 u[8];
                while (true) {
                    bool ok = true;
                    for (int q = 0; q < 2 * nk; ++q) u[q] = ld4<SC>(bx + q * T);
                    for (int q = 0; q < 2 * nk; ++q) ok = ok && u[q].y == seq && u[q].w == seq;
                    if (ok) break;
                }
                for (int k = 0; k < nk; ++k)
                    v[k] = make_double2(__hiloint2double(u[2 * k].z, u[2 * k].x), __hiloint2double(u[2 * k + 1].z, u[2 * k + 1].x));
            };
            if (has_lo) recv(box(kb - 1, 0), GL);
            if (has_hi) recv(box(kb + 1, 1), GR);
            for (int k = 0; k < nk; ++k) {
                if (has_lo) A[k].x += 1e-3 * GL[k].y;
                if (has_hi) D[k].x += 1e-3 * GR[k].y;
            }
        } else {
            double2 *hb = reinterpret_cast<double2 *>(halo);
            auto box = [&](int blk, int dir) { return hb + (((size_t)blk * 2 + dir) * 2 + (seq & 1u)) * 4 * T + tid; };
            auto flag = [&](int blk, int dir) { return flags + (((size_t)blk * 2 + dir) * 2 + (seq & 1u)) * 32 + (VAR == 2 ? warp : 0); };
            if (has_lo)
                for (int k = 0; k < 4; ++k) box(kb, 1)[k * T] = A[k];
            if (has_hi)
                for (int k = 0; k < 4; ++k) box(kb, 0)[k * T] = D[k];
            if (VAR == 2) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    if (has_lo) st_release(flag(kb, 1), seq);
                    if (has_hi) st_release(flag(kb, 0), seq);
                }
            } else {
                __syncthreads();
                if (tid == 0) {
                    __threadfence();
                    if (has_lo) st_release(flag(kb, 1), seq);
                    if (has_hi) st_release(flag(kb, 0), seq);
                }
            }
            for (int i = 0; i < work; ++i) acc = fma(acc, 1.0000001, 1e-9);
            if (has_lo) {
                if (lane == 0) while (ld_acquire(flag(kb - 1, 0)) != seq) {}
                __syncwarp();
                for (int k = 0; k < 4; ++k) GL[k] = __ldcg(box(kb - 1, 0) + k * T);
            }
            if (has_hi) {
                if (lane == 0) while (ld_acquire(flag(kb + 1, 1)) != seq) {}
                __syncwarp();
                for (int k = 0; k < 4; ++k) GR[k] = __ldcg(box(kb + 1, 1) + k * T);
            }
            for (int k = 0; k < 4; ++k) {
                if (has_lo) A[k].x += 1e-3 * GL[k].y;
                if (has_hi) D[k].x += 1e-3 * GR[k].y;
            }
        }
    }
    for (int k = 0; k < 4; ++k) acc += A[k].x + D[k].x;
    sink[(size_t)kb * T + tid] = acc;
}

template <int VAR>
void run(const char *name, int rounds, int work)
{
    const int nblk = 125, T = 512;
    uint4 *halo;
    unsigned *flags;
    double *sink;
    const size_t nh = (size_t)nblk * 2 * 2 * 8 * T;
    cudaMalloc(&halo, nh * sizeof(uint4));
    cudaMalloc(&flags, (size_t)nblk * 2 * 2 * 32 * sizeof(unsigned));
    cudaMalloc(&sink, (size_t)nblk * T * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(halo, 0, nh * sizeof(uint4));
        cudaMemset(flags, 0, (size_t)nblk * 2 * 2 * 32 * sizeof(unsigned));
        void *args[] = {&halo, &flags, &rounds, &work, &sink};
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_xchg<VAR>, dim3(nblk), dim3(T), args, 0, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            printf("%s: launch failed %s\n", name, cudaGetErrorString(e));
            return;
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    printf("%-40s work=%5d : %7.3f us / exchange\n", name, work, 1e3 * best / rounds);
    cudaFree(halo), cudaFree(flags), cudaFree(sink);
}

int main()
{
    const int rounds = 2000;
    for (int work : {0, 500, 2000}) {
        run<0>("LL volatile (2 ch/dir... 64KB+64KB x2)", rounds, work);
        run<1>("LL relaxed.gpu", rounds, work);
        run<2>("plain + per-warp fence/flag", rounds, work);
        run<3>("plain + per-CTA fence/flag", rounds, work);
        run<4>("LL volatile half volume", rounds, work);
    }
    return 0;
}
