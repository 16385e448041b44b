// ionization_b200 -- the ON-CHIP RESIDENT split-operator kernel (sm_100a).
//
// A SphericalHarmonicMesh simulation whose wavefunction fits in the register files of the GPU (l_bound/4 CTAs of
// 4 channels x <= 2048 radial rows: 64 registers of psi per thread, 128 KB per SM; config 3's 500 x 2000 mesh is
// 16 MB on 125 SMs) is advanced by ONE persistent cooperative kernel for a whole stretch of time steps: psi is loaded
// once, stays in registers, and is written back once.  Per time step nothing goes through HBM or L2 except the
// boundary channels that neighbouring CTAs exchange (32 KB per CTA side), so the step is bound by the FP64 pipe and
// by the exchange latency, not by memory.
//
//   CTA k owns channels a,b,c,d = 4k .. 4k+3.  Thread t owns rows 4t .. 4t+3 of all four ("layout 1"), so every
//   l<->l+1 operator of the reference (mesh_operators.py:1037-1080 length gauge, :1204-1408 velocity gauge) on the
//   pairs (a,b), (c,d) [even sweeps] and (b,c) [odd sweeps] is thread-local.  The odd-sweep pairs (4k-1, 4k) and
//   (4k+3, 4k+4) straddle two CTAs: before every odd phase each CTA sends its channels a and d to its neighbours and
//   both CTAs evaluate the straddling pair redundantly (the same scheme as the l-block shards between GPUs).
//
//   Exchange = NCCL-"LL"-style flagged stores: every double travels as two 8-byte {payload32, sequence} words, the
//   receiver spins on its own 128 bytes until all sequence numbers match.  No fence, no barrier, one L2 round trip.
//   The two slots of a mailbox alternate; a sender can never be two exchanges ahead of its receiver because each
//   exchange needs the neighbour's previous one.  All CTAs are co-resident (cooperative launch), every spin has a
//   time-out that raises an abort flag, so a lost message ends the kernel instead of hanging the GPU.
//
//   Crank-Nicolson (evolution_methods.py:98-111 + cy.pyx:9-50) runs on two channels at a time in "layout 2": the two
//   lanes of a lane pair swap half of their rows so that each holds 8 consecutive rows of ONE channel; forward /
//   backward affine recurrences over the 8 rows, Kogge-Stone scan of the chunk maps over the 16 lanes of the same
//   channel, neighbouring-warp inflow through shared memory (same decay-bounded reach as kernels.cuh), transposed
//   back.  The boundary channels (a, d) go first so that their exchange overlaps the solve of (b, c).
//   The LU factors of the CTA's four channels live in shared memory (128 KB) for the whole kernel.
#pragma once
#include "kernels.cuh"

namespace ion {

struct ResidentParams {
    cplx *psi;              // [batch][L][4][T]  internal layout (kernels.cuh)
    const cplx *w;          // [L][4][T]         1/pivot of (1 + i tau H0)
    const double *toff;     // [4][T]            tau * h_off, permuted
    const double *vec;      // [4][T]            rotation coupling vector
    const double *zvec;     // [4][T]            r-pair coupling (velocity gauge)
    const double *zprev;    // [T]
    const double *mask;     // [4][T] or nullptr
    const double *cl;       // [L-1]
    const double *cl2;      // [L-1]
    const double *scal;     // [n_steps][batch]  tau * field
    uint4 *halo;            // LL mailboxes [batch][nblk][2 dirs][2 slots][8][T]
    unsigned *abort_flag;   // set to 1 when an exchange timed out
    long long n_steps;
    long long spin_limit;   // clock64 ticks before a spin gives up
    int L;                  // channels (even)
    int T;                  // threads per CTA = row stride of the layout
    int batch;
    int short_scan;         // reach of the cross-warp inflow (warps); 0 = unbounded
    unsigned seq_base;      // sequence number of the last exchange of the previous launch on these mailboxes
    int dbg;                // experiments only (ION_RES_DBG): 1 = do not wait for the neighbours, 2 = skip CN, 4 = skip the l-sweeps
};

// ---------------------------------------------------------------------------------------------
// LL mailbox
// ---------------------------------------------------------------------------------------------
ION_DEVINL void ll_store(uint4 *p, double v, unsigned flag)
{
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
ION_DEVINL uint4 ll_load(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// box: this thread's column of a mailbox slot, unit u at box[u * T]
ION_DEVINL void ll_send(uint4 *box, int T, const cplx (&v)[4], unsigned seq)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ll_store(box + (2 * k) * T, v[k].x, seq);
        ll_store(box + (2 * k + 1) * T, v[k].y, seq);
    }
}
struct LLState {
    unsigned *abort_flag;
    long long spin_limit;
    bool dead;
};
ION_DEVINL void ll_recv(const uint4 *box, int T, cplx (&v)[4], unsigned seq, LLState &st)
{
    uint4 u[8];
    long long t0 = 0;
    unsigned spins = 0;
    while (true) {
        bool ok = true;
#pragma unroll
        for (int q = 0; q < 8; ++q) u[q] = ll_load(box + q * T);
#pragma unroll
        for (int q = 0; q < 8; ++q) ok = ok && (u[q].y == seq) && (u[q].w == seq);
        if (ok || st.dead) break;
        if ((++spins & 255u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (*(volatile unsigned *)st.abort_flag != 0u) st.dead = true;
            else if (now - t0 > st.spin_limit) {
                atomicExch(st.abort_flag, 1u);
                st.dead = true;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        v[k] = c_make(__hiloint2double((int)u[2 * k].z, (int)u[2 * k].x), __hiloint2double((int)u[2 * k + 1].z, (int)u[2 * k + 1].x));
}

// upper / lower member only of an l-pair rotation (the straddling pairs: the partner belongs to the neighbour CTA)
template <bool REAL>
ION_DEVINL void rotate_upper_only(const cplx (&A)[4], cplx (&B)[4], const RotAngles<4> &ang)  // B = upper member (l+1)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double sn = ang.s[k], cs = ang.c[k];
        const cplx a = A[k], b = B[k];
        if (REAL) B[k] = c_make(fma(cs, b.x, -sn * a.x), fma(cs, b.y, -sn * a.y));
        else B[k] = c_make(fma(cs, b.x, sn * a.y), fma(cs, b.y, -sn * a.x));
    }
}
template <bool REAL>
ION_DEVINL void rotate_lower_only(cplx (&A)[4], const cplx (&B)[4], const RotAngles<4> &ang)  // A = lower member (l)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double sn = ang.s[k], cs = ang.c[k];
        const cplx a = A[k], b = B[k];
        if (REAL) A[k] = c_make(fma(cs, a.x, sn * b.x), fma(cs, a.y, sn * b.y));
        else A[k] = c_make(fma(cs, a.x, sn * b.y), fma(cs, a.y, -sn * b.x));
    }
}

// =============================================================================================
// grid = (ceil(L / 4), batch), block = T (<= 512), cooperative launch.
// VEL = 0: E_e E_o CN E_o E_e mask            
// VEL = 1: h1_e h1_o h2_ee h2_eo h2_oe h2_oo CN h2_oo h2_oe h2_eo h2_ee h1_o h1_e mask   
// The trailing even rotation of step n, the mask and the leading even rotation of step n+1 act on the same pairs and
// are diagonal in r: one rotation by s_n + s_{n+1} (as the streaming path does across kernels).
// =============================================================================================
template <int VEL>
__global__ void __launch_bounds__(512, 1) k_resident(const ResidentParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    const bool odd = (lane & 1) != 0;
    cplx *wsm = reinterpret_cast<cplx *>(smem_raw);  // [2 batches][8 rows][T]   LU factors, layout 2
    cplx *aggsm = wsm + 16 * T;                      // [2 batches][P, Q][T]
    cplx *sm_scan = aggsm + 4 * T;                   // 128 cplx
    double *tosm = reinterpret_cast<double *>(sm_scan + 128);  // [9][T / 2]  tau*off per row of a chunk (+ the row before it)
    cplx *xs = reinterpret_cast<cplx *>(tosm + 9 * (T >> 1) + ((T >> 1) & 1));  // 4 * T cplx (velocity gauge r-pair exchange)

    const int kb = blockIdx.x, nblk = gridDim.x, b = blockIdx.y;
    const int l0 = 4 * kb, L = p.L;
    const bool has_lo = kb > 0, has_hi = (kb + 1 < nblk);
    const bool has_bc = l0 + 2 < L;  // L is even: the last CTA holds either four channels or two
    constexpr bool REAL = (VEL != 0);

    // ---- prologue: LU factors of my four channels into shared memory (layout 2), chunk multipliers ----
    // layout 2: thread (lane pair pp = tid >> 1) holds rows 8 pp .. 8 pp + 7 of channel {a, d}[odd] (batch 0) or {b, c}[odd] (batch 1)
    const int pp = tid >> 1;
    const int TH = T >> 1;
    if (!odd) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tosm[k * TH + pp] = p.toff[(k & 3) * T + 2 * pp + (k >> 2)];
        tosm[8 * TH + pp] = pp > 0 ? p.toff[3 * T + 2 * pp - 1] : 0.0;
    }
#pragma unroll
    for (int bt = 0; bt < 2; ++bt) {
        const int ch = l0 + (bt == 0 ? (odd ? 3 : 0) : (odd ? 2 : 1));
#pragma unroll
        for (int k = 0; k < 8; ++k)
            wsm[(bt * 8 + k) * T + tid] = ch < L ? p.w[((size_t)ch * 4 + (k & 3)) * T + 2 * pp + (k >> 2)] : c_make(1.0, 0.0);
    }
    __syncthreads();
#pragma unroll
    for (int bt = 0; bt < 2; ++bt) {
        const cplx *wcol = wsm + (bt * 8) * T + tid;
        const cplx wprev = tid >= 2 ? wsm[(bt * 8 + 7) * T + tid - 2] : c_zero();
        cplx P = e_of(tosm[8 * TH + pp], wprev), Q = c_make(1.0, 0.0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const cplx e = e_of(tosm[k * TH + pp], wcol[k * T]);
            if (k < 7) P = c_mul(P, e);
            Q = c_mul(Q, e);
        }
        aggsm[(bt * 2 + 0) * T + tid] = P;
        aggsm[(bt * 2 + 1) * T + tid] = Q;
    }
    __syncthreads();

    // ---- psi into registers (layout 1) ----
    cplx A[4], B[4], C[4], D[4];
    {
        cplx *base = p.psi + ((size_t)b * L + l0) * 4 * T;
        load_rows<4>(A, base, T, tid, l0 < L);
        load_rows<4>(B, base + 4 * T, T, tid, l0 + 1 < L);
        load_rows<4>(C, base + 8 * T, T, tid, l0 + 2 < L);
        load_rows<4>(D, base + 12 * T, T, tid, l0 + 3 < L);
    }
    auto clv = [&](const double *cl, int l) -> double { return (l >= 0 && l + 1 < L) ? cl[l] : 0.0; };

    // mailboxes: [b][blk][dir][slot][8][T]; dir 0 = sent upwards (my channel d), dir 1 = sent downwards (my channel a)
    const size_t box_sz = (size_t)8 * T;
    auto box = [&](int blk, int dir, unsigned seq) -> uint4 * {
        return p.halo + ((((size_t)b * nblk + blk) * 2 + dir) * 2 + (seq & 1u)) * box_sz + tid;
    };
    LLState ll;
    ll.abort_flag = p.abort_flag;
    ll.spin_limit = p.spin_limit;
    ll.dead = (p.dbg & 1) != 0;
    unsigned seq = p.seq_base;

    auto send_edges = [&]() {
        ++seq;
        if (has_lo) ll_send(box(kb, 1, seq), T, A, seq);
        if (has_hi) ll_send(box(kb, 0, seq), T, D, seq);
    };
    cplx GL[4], GR[4];
    auto recv_edges = [&]() {
        if (has_lo) ll_recv(box(kb - 1, 0, seq), T, GL, seq, ll);
        if (has_hi) ll_recv(box(kb + 1, 1, seq), T, GR, seq, ll);
    };
    auto load4 = [&](double (&v)[4], const double *src) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = src[k * T + tid];
    };
    auto cn_batch = [&](cplx (&X)[4], cplx (&Y)[4], int bt) {
        if (p.dbg & 2) return;
        cplx Z[8];
        pair_transpose_in(X, Y, Z, odd);
        const cplx wprev = tid >= 2 ? wsm[(bt * 8 + 7) * T + tid - 2] : c_zero();
        cn8(Z, wsm + (bt * 8) * T + tid, T, tosm + pp, wprev, aggsm[(bt * 2 + 0) * T + tid], aggsm[(bt * 2 + 1) * T + tid], tid, T, sm_scan,
            p.short_scan);
        pair_transpose_out(Z, X, Y, odd);
    };
    // l-pair rotation sweep on the odd pairs: (b, c) local, then the straddling pairs with the neighbours' channels
    auto odd_rotation = [&](double s) {
        if (p.dbg & 4) {
            recv_edges();
            return;
        }
        double vec[4];
        load4(vec, p.vec);
        if (has_bc) {
            const RotAngles<4> ang = rot_angles<4>(vec, s * clv(p.cl, l0 + 1));
            rotate_pair<4, REAL>(B, C, ang);
        }
        recv_edges();
        if (has_lo) {
            const RotAngles<4> ang = rot_angles<4>(vec, s * clv(p.cl, l0 - 1));
            rotate_upper_only<REAL>(GL, A, ang);
        }
        if (has_hi) {
            const RotAngles<4> ang = rot_angles<4>(vec, s * clv(p.cl, l0 + 3));
            rotate_lower_only<REAL>(D, GR, ang);
        }
    };
    auto even_rotation = [&](double s, bool with_mask) {
        if (p.dbg & 4) return;
        double vec[4];
        load4(vec, p.vec);
        {
            const RotAngles<4> ang = rot_angles<4>(vec, s * clv(p.cl, l0));
            rotate_pair<4, REAL>(A, B, ang);
        }
        if (has_bc) {
            const RotAngles<4> ang = rot_angles<4>(vec, s * clv(p.cl, l0 + 2));
            rotate_pair<4, REAL>(C, D, ang);
        }
        if (with_mask && p.mask) {
            double mk[4];
            load4(mk, p.mask);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                A[k] = c_scale(A[k], mk[k]);
                B[k] = c_scale(B[k], mk[k]);
                C[k] = c_scale(C[k], mk[k]);
                D[k] = c_scale(D[k], mk[k]);
            }
        }
    };
    // velocity-gauge h2 bricks on the even pairs / the odd pairs
    auto h2_even = [&](double s, bool reverse) {
        double zv[4];
        load4(zv, p.zvec);
        const double zp = p.zprev[tid];
        {
            const RPairAngles<4> ang = rpair_angles<4>(zv, zp, s * clv(p.cl2, l0));
            h2_pair<4>(A, B, ang, reverse, tid, T, xs);
        }
        if (has_bc) {
            const RPairAngles<4> ang = rpair_angles<4>(zv, zp, s * clv(p.cl2, l0 + 2));
            h2_pair<4>(C, D, ang, reverse, tid, T, xs);
        }
    };
    auto h2_odd = [&](double s, bool reverse) {
        double zv[4];
        load4(zv, p.zvec);
        const double zp = p.zprev[tid];
        if (has_bc) {
            const RPairAngles<4> ang = rpair_angles<4>(zv, zp, s * clv(p.cl2, l0 + 1));
            h2_pair<4>(B, C, ang, reverse, tid, T, xs);
        }
        recv_edges();
        if (has_lo) {  // CTA-uniform branches: the barriers inside h2_pair are reached by every thread
            const RPairAngles<4> ang = rpair_angles<4>(zv, zp, s * clv(p.cl2, l0 - 1));
            h2_pair<4>(GL, A, ang, reverse, tid, T, xs);
        }
        if (has_hi) {
            const RPairAngles<4> ang = rpair_angles<4>(zv, zp, s * clv(p.cl2, l0 + 3));
            h2_pair<4>(D, GR, ang, reverse, tid, T, xs);
        }
    };

    double s_prev = 0.0;
    for (long long n = 0; n < p.n_steps; ++n) {
        const double s = p.scal[(size_t)n * p.batch + b];
        even_rotation(s_prev + s, n > 0);  // E_e / h1_e of this step fused with the previous step's tail and mask
        send_edges();
        odd_rotation(s);                   // E_o / h1_o
        if (VEL) {
            h2_even(s, false);             // h2_ee, h2_eo
            send_edges();
            h2_odd(s, false);              // h2_oe, h2_oo
        }
        cn_batch(A, D, 0);                 // boundary channels first: their exchange overlaps the solve of (b, c)
        send_edges();
        cn_batch(B, C, 1);
        if (VEL) {
            h2_odd(s, true);               // h2_oo, h2_oe
            h2_even(s, true);              // h2_eo, h2_ee
            send_edges();
        }
        odd_rotation(s);                   // E_o / h1_o
        s_prev = s;
    }
    even_rotation(s_prev, true);           // tail of the last step + mask

    {
        cplx *base = p.psi + ((size_t)b * L + l0) * 4 * T;
        store_rows<4>(A, base, T, tid, l0 < L);
        store_rows<4>(B, base + 4 * T, T, tid, l0 + 1 < L);
        store_rows<4>(C, base + 8 * T, T, tid, l0 + 2 < L);
        store_rows<4>(D, base + 12 * T, T, tid, l0 + 3 < L);
    }
}

}  // namespace ion
