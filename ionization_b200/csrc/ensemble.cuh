// ionization_b200 -- the folded length-gauge time step for SCAN ENSEMBLES (sm_100a): persistent CTAs with a prefetch pipeline.
//
// k_unit<PROG_LEN_STEP> (kernels.cuh) runs one CTA per (odd channel pair, ensemble member): load psi (own two channels and the two
// read-only even-pair partners), rotate, Crank-Nicolson, rotate, store.  For an ensemble that is hundreds of waves of short-lived
// CTAs whose DRAM latency is exposed in every one of them (ncu: 32 % long-scoreboard stalls, two 256-thread CTAs per SM).
// Here 2 CTAs per SM stay resident and walk over the (member, pair) tasks MEMBER-MAJOR with a grid stride -- neighbouring
// pairs of the same member are in flight at the same time on other SMs, so the read-only partner channels still hit L2 -- and
//   * the psi of the NEXT task (4 channels x 4 rows per thread = 64 KB per CTA at T = 256) is prefetched into shared memory while
//     the current task computes -- by the TMA engine (BULK: four consecutive channels are ONE contiguous 64 KB block of the
//     row-interleaved layout; cp.async.bulk issued by one thread behind the forward scan's barrier, completion on an mbarrier), or,
//     for A/B timing, with 16 cp.async per thread (every thread stages exactly the values it will read itself);
//   * the LU factors of the current pair are staged with cp.async at the top of the task and arrive under the trigonometry;
//   * tau * h_off of the rows (the same for every pair) is staged once per CTA.
// The arithmetic is the layout-2 path of k_unit<PROG_LEN_STEP>, instruction for instruction (same helpers): results are identical.
// The two single channels of an odd sweep (l = 0 and l = L - 1) are left to k_unit (UnitParams.unit0 / unit_stride).
#pragma once
#include "kernels.cuh"

namespace ion {

constexpr int ENS_T = 256;  // threads per CTA = threads per channel: r_points in (896, 1024].  Measured on smaller meshes (128- and
                            // 64-thread channels, where k_unit already keeps 4 to 8 CTAs per SM in different phases): 40 % SLOWER
                            // than one CTA per task (profiles/r01c_whatif_experiments.md), so the kernel is used for T = 256 only.

// dynamic shared memory: scan scratch 256 cplx | LU factors 8*T cplx | psi stage 16*T cplx | tau*off 9*(T/2) doubles
inline size_t ens_smem_bytes(int T) { return (256 + 8 * (size_t)T + 16 * (size_t)T) * sizeof(cplx) + 9 * (size_t)(T / 2) * sizeof(double) + 16; }

// TASK ORDER.  An item is (block of ENS_MB consecutive members, channel pair); items are numbered block-major, pair-minor, and a
// CTA takes the items blockIdx.x, blockIdx.x + gridDim.x, ... and walks through the members of an item one after the other.
// So (1) the CTAs in flight work on neighbouring pairs of the same members at the same time -- the read-only partner channels
// hit L2 as before -- and (2) consecutive tasks of a CTA share their pair: the LU factors stay staged in shared memory and the
// pair's scan multipliers hit L1 for ENS_MB - 1 of every ENS_MB tasks (one third less L2 -> SM traffic, no factor wait).
#ifndef ION_ENS_MB
#define ION_ENS_MB 16
#endif
constexpr int ENS_MB = ION_ENS_MB;

// BULK: the psi of the next task -- four consecutive channels, i.e. ONE contiguous 64 KB block in the row-interleaved layout -- is moved
// by the TMA engine (cp.async.bulk, four 16 KB copies issued by one thread, completion on an mbarrier) instead of 16 cp.async per thread;
// the copy is issued behind the forward scan's barrier of the current task, when every thread has read the staging buffer.
template <bool BULK>
__global__ void __launch_bounds__(ENS_T, 2) k_len_ens(const UnitParams p, int n_pairs, int batch)
{
    constexpr int M = 4, T = ENS_T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm_scan = reinterpret_cast<cplx *>(smem_raw);
    cplx *wsm = sm_scan + 256;                               // [8][T]
    cplx *stage = wsm + 8 * T;                               // [4 channels][4 rows][T]
    double *tosm = reinterpret_cast<double *>(stage + 16 * T);  // [9][T/2]
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(tosm + 9 * (T / 2));  // BULK: psi of the next task has landed

    const int tl = threadIdx.x, pp = tl >> 1, TH = T >> 1;
    const bool odd = (tl & 1) != 0;
    const size_t chan = (size_t)M * T;
    pdl_launch_dependents();

    // ---- once per CTA: everything that does not depend on the pair or the member ----
    if (!odd) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tosm[k * TH + pp] = p.toff[(k & 3) * T + 2 * pp + (k >> 2)];
        tosm[8 * TH + pp] = p.toff_prev[2 * pp];
    }
    const bool masked = (p.flags & F_MASK) != 0;

    const int n_blocks = (batch + ENS_MB - 1) / ENS_MB;
    const long long n_items = (long long)n_pairs * n_blocks;
    auto member_of = [&](long long item, int mem) { return (int)(item / n_pairs) * ENS_MB + mem; };
    auto prefetch_psi = [&](long long item, int mem) {
        const long long b = member_of(item, mem);
        const int l0 = 2 * (int)(item % n_pairs) + 1;
        if constexpr (BULK) {
            if (tl == 0) {
                const cplx *src = p.psi + ((size_t)b * p.L + (l0 - 1)) * chan;
                mbar_expect_tx(mbar, (unsigned)(4 * chan * sizeof(cplx)));
#pragma unroll
                for (int c = 0; c < 4; ++c) bulk_g2s(stage + (size_t)c * chan, src + (size_t)c * chan, (unsigned)(chan * sizeof(cplx)), mbar);
            }
        } else {
            const cplx *src = p.psi + ((size_t)b * p.L + (l0 - 1)) * chan + tl;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int k = 0; k < 4; ++k) cp_async16(stage + (size_t)(c * 4 + k) * T + tl, src + (size_t)c * chan + (size_t)k * T);
            }
        }
    };

    if constexpr (BULK) {
        if (tl == 0) mbar_init(mbar, 1);
        __syncthreads();
    }
    unsigned phase = 0;
    pdl_wait();
    long long item = blockIdx.x;
    int mem = 0;
    if (item < n_items) prefetch_psi(item, 0);
    cp_async_commit();  // group P(task)

    while (item < n_items) {
        const long long b = member_of(item, mem);
        const int l0 = 2 * (int)(item % n_pairs) + 1;  // the pair (l0, l0 + 1); partners l0 - 1 and l0 + 2
        // the task after this one: the next member of the item, else the first member of the CTA's next item
        long long next_item = item;
        int next_mem = mem + 1;
        if (next_mem >= ENS_MB || member_of(item, next_mem) >= batch) next_item = item + gridDim.x, next_mem = 0;
        // ---- LU factors of this pair -> shared memory, permuted into the layout-2 order (as k_unit); once per item ----
        if (mem == 0) {
            const cplx *w0 = p.w + (size_t)l0 * chan + tl;
            cplx *dst = wsm + (size_t)(4 * (tl & 1)) * T + (tl & ~1);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                cp_async16(dst + (size_t)k4 * T, w0 + (size_t)k4 * T);
                cp_async16(dst + (size_t)k4 * T + 1, w0 + chan + (size_t)k4 * T);
            }
        }
        cp_async_commit();  // group F(task) (empty when the factors are already staged)
        cplx P8, Q8, wprev;
        {
            const size_t ch = (size_t)(l0 + (odd ? 1 : 0)) * T + 2 * pp;
            P8 = c_mul(ld_c(p.aggP + ch), ld_c(p.aggP + ch + 1));
            Q8 = c_mul(ld_c(p.aggQ + ch), ld_c(p.aggQ + ch + 1));
            wprev = pp > 0 ? ld_c(p.w + (size_t)(l0 + (odd ? 1 : 0)) * chan + 3 * (size_t)T + 2 * pp - 1) : c_zero();
        }
        const double sa = p.scal_a ? p.scal_a[b] : 0.0, sb = p.scal_b ? p.scal_b[b] : 0.0;
        double vec[M];  // re-read per task (L1): keeping it across tasks costs registers the solve needs
        load_vec<M>(vec, p.vec, T, tl, true);
        const RotAngles<M> rang = rot_angles_auto<M>(vec, p.vec_dv, sa * p.cl[l0]);
        cplx A[M], B[M];
        {
            const RotAngles<M> eangA = rot_angles_auto<M>(vec, p.vec_dv, (sa + sb) * p.cl[l0 - 1]);
            const RotAngles<M> eangB = rot_angles_auto<M>(vec, p.vec_dv, (sa + sb) * p.cl[l0 + 1]);
            // ---- psi of this task: staged by this very thread during the previous task ----
            if constexpr (BULK) {
                mbar_wait(mbar, phase);
                phase ^= 1u;
            } else {
                asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but F(task) has landed
            }
            cplx QA[M], QB[M];
#pragma unroll
            for (int k = 0; k < M; ++k) {
                QA[k] = stage[(size_t)(0 * 4 + k) * T + tl];
                A[k] = stage[(size_t)(1 * 4 + k) * T + tl];
                B[k] = stage[(size_t)(2 * 4 + k) * T + tl];
                QB[k] = stage[(size_t)(3 * 4 + k) * T + tl];
            }
            if (!BULK && next_item < n_items) prefetch_psi(next_item, next_mem);
            cp_async_commit();  // group P(next) (possibly empty)
            rotate_member<M>(A, QA, eangA);
            rotate_member<M>(B, QB, eangB);
        }
        if (masked) {
            double mk[M];
            load_vec<M>(mk, p.mask, T, tl, true);
#pragma unroll
            for (int k = 0; k < M; ++k) {
                A[k] = c_scale(A[k], mk[k]);
                B[k] = c_scale(B[k], mk[k]);
            }
        }
        rotate_pair<M, false>(A, B, rang);
        if constexpr (BULK) asm volatile("cp.async.wait_group 0;" ::: "memory");  // F(task) has landed
        else asm volatile("cp.async.wait_group 1;" ::: "memory");                  // F(task) has landed; P(next) may still be in flight
        __syncwarp();  // a thread reads factors staged by itself and by its lane-pair partner only
        {
            cplx Z[8];
            pair_transpose_in(A, B, Z, odd);
            auto hook = [&]() {
                if (BULK && next_item < n_items) prefetch_psi(next_item, next_mem);
            };
            cn8(Z, wsm + tl, T, tosm + pp, wprev, P8, Q8, tl, T, sm_scan, p.short_scan, hook);
            pair_transpose_out(Z, A, B, odd);  // its shuffles also order the pair's reads of wsm before the next task's staging
        }
        rotate_pair<M, false>(A, B, rang);
        cplx *obase = p.psi_out + ((size_t)b * p.L + l0) * chan;
        store_rows<M>(A, obase, T, tl, true);
        store_rows<M>(B, obase + chan, T, tl, true);
        item = next_item;
        mem = next_mem;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace ion
