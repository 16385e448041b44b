// ionization_b200 -- halo exchange between l-block shards over NVLink peer memory (sm_100a).
//
// One process per GPU owns one l-block of a large SphericalHarmonicMesh simulation.  Shards are cut at
// even channels, so only the odd-parity kernels touch a pair that straddles a cut; each shard keeps one ghost channel
// per neighbour and both shards evaluate the straddling pair redundantly.  Before every odd-parity kernel the ghost
// must hold the neighbour's current boundary channel.  This kernel does that exchange with plain stores into the
// neighbour's memory (CUDA IPC mapping of its "halo block": flags + two staging slots per side) -- no host involvement,
// no NCCL call, capturable into the step loop:
//
//   k = exchanges completed so far + 1                       (device-side counter: the kernel takes no per-call argument)
//   1. store my boundary channel into the neighbour's staging slot k & 1 (16-byte stores over NVLink)
//   2. system-scope fence, then ARRIVE:                                                  -> its  arrive[side'] = k
//   3. wait until the neighbour's data has arrived here                                  my   arrive[side]  >= k
//   4. copy my staging slot k & 1 into my ghost channel (local)
//
// Two slots make a "ready to receive" hand-shake unnecessary: the neighbour can only be storing exchange k into slot
// k & 1 after it has seen my ARRIVE k-1, i.e. after I launched exchange k-1, which in stream order follows my step 4 of
// exchange k-2 -- the last reader of that slot.  One NVLink latency + one fence round trip per exchange.
// The wait times out into an abort flag instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace ion {

struct HaloParams {
    unsigned long long *flags;          // my flag block [HF_COUNT]
    unsigned long long *peer_flags[2];  // the neighbours' flag blocks (nullptr: no neighbour on that side)
    const cplx *src[2];                 // my boundary channel towards side 0 (lower) / 1 (upper)
    cplx *peer_stage[2];                // the neighbour's two staging slots for the side facing me (peer memory)
    const cplx *my_stage[2];            // my two staging slots for side 0 / 1
    cplx *ghost[2];                     // my ghost channels
    long long n;                        // complex values per channel (Rp * batch)
    long long spin_limit;               // clock64 ticks before a wait gives up
};

// grid = (n_ctas, 2 sides), block = 256
__global__ void __launch_bounds__(256) k_halo_exchange(const HaloParams p)
{
    __shared__ int ok_sm;
    const int side = blockIdx.y, other = 1 - side;
    unsigned long long *flags = p.flags;
    const unsigned long long k = ld_acquire_sys(flags + HF_SEQ) + 1ull;
    const long long slot = (long long)(k & 1ull) * p.n;
    const bool have = p.peer_flags[side] != nullptr;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    if (have) {
        const cplx *src = p.src[side];
        cplx *dst = p.peer_stage[side] + slot;
        for (long long i = i0; i < p.n; i += stride) dst[i] = src[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long done = atomicAdd(flags + HF_DONE_SIDE + side, 1ull);
            if (done == gridDim.x - 1) {  // everything of this side is on its way and fenced: publish
                flags[HF_DONE_SIDE + side] = 0ull;
                __threadfence_system();
                st_release_sys(p.peer_flags[side] + HF_ARRIVE + other, k);
            }
            ok_sm = halo_spin(flags + HF_ARRIVE + side, k, flags, p.spin_limit) ? 1 : 0;
            if (!ok_sm) {  // tell both neighbours: they must not keep stepping with a ghost channel that never arrived here
                for (int q = 0; q < 2; ++q)
                    if (p.peer_flags[q]) st_release_sys(p.peer_flags[q] + HF_ABORT, 1ull);
            }
        }
        __syncthreads();
        if (ok_sm) {
            const cplx *st = p.my_stage[side] + slot;
            cplx *g = p.ghost[side];
            for (long long i = i0; i < p.n; i += stride) g[i] = st[i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long all = atomicAdd(flags + HF_DONE, 1ull);
        if (all == (unsigned long long)gridDim.x * gridDim.y - 1ull) {
            flags[HF_DONE] = 0ull;
            __threadfence();
            st_release_sys(flags + HF_SEQ, k);
        }
    }
}

}  // namespace ion
