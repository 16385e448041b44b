// ionization_b200 -- the hot-path kernels (sm_100a).
//
// DATA LAYOUT:  one l-channel of R radial points is owned by a CTA of T threads, thread t
// holding the M CONSECUTIVE rows i = t*M + k, k = 0..M-1, in registers.  So that every global access is a
// fully coalesced 128-bit access, each channel is stored "row-interleaved":
//
//        position(i) = (i % M) * T + (i / M)            i.e. psi[sim][l][k][t],  Rp = M*T >= R, zero padded.
//
// The same permutation is applied to every per-r coefficient vector, so kernels that are point-wise in r
// never need to know it; kernels that couple r-neighbours (Crank-Nicolson, the velocity-gauge r-pair
// rotations) find the neighbours i-1 / i+1 in the same thread, or in thread t-1 / t+1 at chunk edges.
//
// Every kernel works on "units": a unit is a pair of channels (l, l+1) with l % 2 == parity, or a single
// left-over channel (l = 0 in odd sweeps, the last channel when it has no partner).  All the operators of the
// reference's split-operator step are local to such a pair, so each kernel streams psi
// through the SM exactly once: 128-bit loads -> registers -> 128-bit stores.
#pragma once
#include "common.cuh"

namespace ion {

enum : int {
    F_MASK = 1,       // multiply by mask[r] at the end (mesh/meshes.py:257)
    F_REAL_ROT = 2,   // rotation [[c, s], [-s, c]] (velocity gauge h1 / line) instead of [[c, -is], [-is, c]]
    F_H2_REVERSE = 4, // r-sublayer order: default (r-even, r-odd); reversed (r-odd, r-even)
    F_SOLVE_ONLY = 8, // PROG_CN: (1 + i tau H0)^-1 g instead of the Crank-Nicolson (1 + i tau H0)^-1 (1 - i tau H0) g  (ADI, adi.cuh)
};

struct UnitParams {
    cplx *psi;               // [batch][L][M][T]
    cplx *psi_out;           // where the results go: psi itself except for the out-of-place PROG_LEN_STEP
    const cplx *w;           // [L][M][T]   1/pivot of (1 + i tau H0), permuted      (k_factor)
    const cplx *aggP;        // [L][T]      forward chunk multipliers
    const cplx *aggQ;        // [L][T]      backward chunk multipliers
    const double *toff;      // [M][T]      tau * h_off[i]   (0 for i >= R-1)
    const double *toff_prev; // [T]         tau * h_off[t*M - 1] (0 for t = 0)
    const double *vec;       // [M][T]      rotation coupling vector (x_j or y_j), 0 in the padding
    const double *zvec;      // [M][T]      r-pair coupling z_j (0 for i >= R-1)
    const double *zprev;     // [T]         z at row t*M - 1
    const double *mask;      // [M][T] or nullptr
    const double *cl;        // [L_total-1] l-pair coefficient for rotations (c_l or c_l*(l+1)), GLOBAL l
    const double *cl2;       // [L_total-1] l-pair coefficient for the r-pair rotations
    const cplx *th;          // [M][T]      tau * h_diag of a LineMesh (time-dependent Crank-Nicolson, PROG_LINE_CN)
    const double *scal_a;    // [batch] per-simulation scalar s = tau*field of this step (or nullptr = 0)
    const double *scal_b;    // [batch] second scalar fused in (next step's), or nullptr
    int L;                   // channels held in psi (owned + ghost channels of an l-block shard)
    int T;                   // TT: threads per channel = row stride of the interleaved layout (Rp = M * TT)
    int S;                   // r-segments per channel (1: the CTA covers the whole channel)
    int T_seg;               // interior threads of a segment
    int H;                   // halo threads on each side of a segment (0 when S == 1)
    int l_begin;             // global index of channel 0 of psi
    int parity;              // parity (in GLOBAL l) of the lower channel of a pair
    int flags;
    int short_scan;          // reach of the cross-warp inflow of the CN scans in warps (0: full scan; see common.cuh)
    // fused observation (PROG_LEN_STEP_OBS): the state after the previous step's mask is reduced inside this step's kernel
    const double *obs_rvec;       // [M][T] permuted r_j
    const cplx *obs_state_rows;   // [n_states][M][T]
    const int *obs_state_first;   // [L + 1]
    const int *obs_state_order;   // [n_states]
    double *obs_partial;          // [batch][L][4 + n_radii]   (k_observe's layout)
    double *obs_ip;               // [batch][n_states][2]
    double obs_radii[8];
    double obs_ipm;
    int obs_n_radii, obs_n_states;
    unsigned obs_what;
    double vec_dv;           // != 0: vec is linear in the row index with this increment per row (length gauge on the uniform radial grid)
    int unit0, unit_stride;  // unit of CTA x = unit0 + x * unit_stride (0, 1: all units; the ensemble kernel leaves the single channels to k_unit)
    // PROG_LEN_STEP_HALO: the folded step of an l-block shard cut at odd channels with the halo exchange fused in (see k_unit)
    unsigned long long *hf_flags;          // my flag block (common.cuh: HF_*)
    unsigned long long *hf_peer_flags[2];  // the neighbours' flag blocks (peer memory)
    cplx *hf_peer_fstage[2];               // the neighbour's two fused staging slots for the side facing me (peer memory)
    const cplx *hf_my_fstage[2];           // my two fused staging slots for side 0 (lower) / 1 (upper)
    unsigned long long *hf_sent;           // [2][S]: launches so far, kept by the boundary CTA (side, segment)
    long long hf_spin_limit;
    int hf_unit[2];                        // the unit whose even-pair partner is the lower / upper ghost channel (-1: no neighbour)
    int hf_consume;                        // the ghost channels of this launch are the previous launch's deliveries (else: already in psi)
    int n_units, boundary_first;           // boundary_first: CTA order (first unit, last unit, interior units) instead of ascending
};

// unit -> (first local channel, is pair).  Pairs are (l, l+1) with global l % 2 == parity.
ION_DEVINL void unit_channels(const UnitParams &p, int unit, int &l0, bool &pair)
{
    // local parity of pair starts
    int lp = (p.parity - (p.l_begin & 1)) & 1;
    if (lp == 0) {
        l0 = 2 * unit;
    } else {
        l0 = unit == 0 ? 0 : 2 * unit - 1;
        if (unit == 0) {
            pair = false;
            return;
        }
    }
    pair = (l0 + 1 < p.L);
}
inline int num_units(int L, int l_begin, int parity)
{
    int lp = (parity - (l_begin & 1)) & 1;
    return lp == 0 ? (L + 1) / 2 : 1 + L / 2;
}

// `ok`: the thread maps to an existing row chunk of the channel (halo threads beyond either end of it do not)
template <int M>
ION_DEVINL void load_rows(cplx (&g)[M], const cplx *base, int T, int t, bool ok)
{
#pragma unroll
    for (int k = 0; k < M; ++k) g[k] = ok ? ld_c(base + k * T + t) : c_zero();
}
// the same through L2 only (data another GPU stored into this one's memory while the kernel may already be running)
template <int M>
ION_DEVINL void load_rows_cg(cplx (&g)[M], const cplx *base, int T, int t, bool ok)
{
#pragma unroll
    for (int k = 0; k < M; ++k) g[k] = ok ? __ldcg(reinterpret_cast<const double2 *>(base + k * T + t)) : c_zero();
}
template <int M>
ION_DEVINL void store_rows(const cplx (&g)[M], cplx *base, int T, int t, bool ok)
{
    if (!ok) return;
#pragma unroll
    for (int k = 0; k < M; ++k) st_c(base + k * T + t, g[k]);
}
template <int M>
ION_DEVINL void load_vec(double (&v)[M], const double *base, int T, int t, bool ok)
{
#pragma unroll
    for (int k = 0; k < M; ++k) v[k] = ok ? base[k * T + t] : 0.0;
}

// ---------------------------------------------------------------------------------------------
// l<->l+1 rotation of one pair, all rows of this thread.
//   complex: [[cos a, -i sin a], [-i sin a, cos a]]   mesh_operators.py:1055-1075
//   real:    [[cos a,  sin a], [-sin a,  cos a]]      mesh_operators.py:1204-1245
// ---------------------------------------------------------------------------------------------
template <int M>
struct RotAngles {
    double c[M], s[M];
};
template <int M>
ION_DEVINL RotAngles<M> rot_angles(const double (&vec)[M], double sc)
{
    RotAngles<M> a;
    double th[M];
#pragma unroll
    for (int k = 0; k < M; ++k) th[k] = sc * vec[k];
    fast_sincos_n<M>(th, a.s, a.c);
    return a;
}
// The same for a coupling vector that is LINEAR in the row index, vec[i] = vec[0] + i * dv (the length gauge: x_j = -q r_j on
// the uniform radial grid): theta_k = sc * (v0 + k dv), so one sincos for the thread's first row, one for the increment
// (the same for every thread of the pair) and M - 1 complex multiplications replace M sincos evaluations.  Rounding: each
// multiplication adds <= 2 ulp to cos/sin, i.e. < 1e-15 absolute after M - 1 = 3 of them.
template <int M>
ION_DEVINL RotAngles<M> rot_angles_linear(double v0, double dv, double sc)
{
    RotAngles<M> a;
    const double th[2] = {sc * v0, sc * dv};
    double sn[2], cs[2];
    fast_sincos_n<2>(th, sn, cs);
    a.c[0] = cs[0];
    a.s[0] = sn[0];
#pragma unroll
    for (int k = 1; k < M; ++k) {
        a.c[k] = fma(a.c[k - 1], cs[1], -a.s[k - 1] * sn[1]);
        a.s[k] = fma(a.s[k - 1], cs[1], a.c[k - 1] * sn[1]);
    }
    return a;
}
// vec_dv != 0: the host verified (ion_sim_set_len_coupling) that the coupling vector is linear in the row index to rounding
template <int M>
ION_DEVINL RotAngles<M> rot_angles_auto(const double (&vec)[M], double vec_dv, double sc)
{
    return vec_dv != 0.0 ? rot_angles_linear<M>(vec[0], vec_dv, sc) : rot_angles<M>(vec, sc);
}
template <int M, bool REAL>
ION_DEVINL void rotate_pair(cplx (&A)[M], cplx (&B)[M], const RotAngles<M> &ang)
{
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const double sn = ang.s[k], cs = ang.c[k];
        cplx a = A[k], b = B[k];
        if (REAL) {
            A[k] = c_make(fma(cs, a.x, sn * b.x), fma(cs, a.y, sn * b.y));
            B[k] = c_make(fma(cs, b.x, -sn * a.x), fma(cs, b.y, -sn * a.y));
        } else {
            A[k] = c_make(fma(cs, a.x, sn * b.y), fma(cs, a.y, -sn * b.x));
            B[k] = c_make(fma(cs, b.x, sn * a.y), fma(cs, b.y, -sn * a.x));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Crank-Nicolson on one channel held in registers:  g <- 2 (1 + i tau H0)^-1 g - g
//   ( = (1 + i tau H0)^-1 (1 - i tau H0) g ; evolution_methods.py:98-111 + cy.pyx:9-50 )
// with the LU factors precomputed (k_factor): forward  y_i = g_i + e_{i-1} y_{i-1},
// backward x_i = w_i y_i + e_i x_{i+1},  e_i = -i (tau off_i) w_i.
// Each thread runs the two recurrences over its M rows twice: once with a zero inflow to get its chunk's
// affine map, then -- after a block-wide scan of those maps -- with the true inflow.
// sm: 4*32 cplx of shared memory private to this call.
// ---------------------------------------------------------------------------------------------
// LU factors of one channel for this thread's rows, loaded up-front so that their latency overlaps whatever
// precedes the Crank-Nicolson solve in the kernel (rotations, r-pair bricks, the other channel's solve).
template <int M>
struct CnFactors {
    cplx w[M];   // 1 / pivot
    cplx wprev;  // w of row t*M - 1 (thread t-1's last row)
    cplx P, Q;   // chunk aggregates
};
template <int M>
ION_DEVINL void cn_load(CnFactors<M> &f, const cplx *__restrict__ wch, const cplx *__restrict__ aggP,
                        const cplx *__restrict__ aggQ, int t, int T, bool ok)
{
#pragma unroll
    for (int k = 0; k < M; ++k) f.w[k] = ok ? ld_c(wch + k * T + t) : c_make(1.0, 0.0);
    f.wprev = (ok && t > 0) ? ld_c(wch + (M - 1) * T + t - 1) : c_zero();
    f.P = ok ? ld_c(aggP + t) : c_zero();
    f.Q = ok ? ld_c(aggQ + t) : c_zero();
}

template <int M>
ION_DEVINL void cn_channel(cplx (&g)[M], const CnFactors<M> &f, const double (&toff)[M], double toff_prev, int t, int T,
                           cplx *sm, int short_scan)
{
    const cplx(&w)[M] = f.w;
    const cplx Pt = f.P, Qt = f.Q;
    cplx e[M];
#pragma unroll
    for (int k = 0; k < M; ++k) e[k] = c_make(toff[k] * w[k].y, -toff[k] * w[k].x);
    const cplx elink = c_make(toff_prev * f.wprev.y, -toff_prev * f.wprev.x);
    // forward, zero inflow
    cplx z = g[0];
#pragma unroll
    for (int k = 1; k < M; ++k) z = c_fma(e[k - 1], z, g[k]);
    cplx yin = affine_scan_block_exclusive<true>(Pt, z, sm, sm + 32, t, T, short_scan);
    // forward, true inflow; u = w*y
    cplx u[M];
    cplx y = c_fma(elink, yin, g[0]);
    u[0] = c_mul(w[0], y);
#pragma unroll
    for (int k = 1; k < M; ++k) {
        y = c_fma(e[k - 1], y, g[k]);
        u[k] = c_mul(w[k], y);
    }
    // backward, zero inflow
    z = u[M - 1];
#pragma unroll
    for (int k = M - 2; k >= 0; --k) z = c_fma(e[k], z, u[k]);
    cplx xin = affine_scan_block_exclusive<false>(Qt, z, sm + 64, sm + 96, t, T, short_scan);
    // backward, true inflow; out = 2x - g
    cplx x = c_fma(e[M - 1], xin, u[M - 1]);
    g[M - 1] = c_make(fma(2.0, x.x, -g[M - 1].x), fma(2.0, x.y, -g[M - 1].y));
#pragma unroll
    for (int k = M - 2; k >= 0; --k) {
        x = c_fma(e[k], x, u[k]);
        g[k] = c_make(fma(2.0, x.x, -g[k].x), fma(2.0, x.y, -g[k].y));
    }
}

// =============================================================================================
// "Layout 2" Crank-Nicolson for a PAIR of channels.  In layout 1 a thread holds rows 4t .. 4t+3 of both channels of
// its pair and would run four scans (two per channel) over 32 lanes.  For the solve the two lanes of a lane pair swap
// half of their rows (one shuffle per value), so that the even lane holds rows 8p .. 8p+7 of the lower channel and the
// odd lane the same rows of the upper channel: one solve per thread over twice the rows, scans over the 16 lanes of
// the same channel (4 Kogge-Stone steps instead of 5), both channels in the same instruction stream, half the barriers.
// Per point this is ~27 % fewer FP64 instructions and ~40 % fewer shuffles than two layout-1 solves.  The LU factors
// are staged in shared memory (cp.async during the psi-independent prologue) in the order the solve reads them.
// =============================================================================================
// ---------------------------------------------------------------------------------------------
// Affine scan over the lanes of the same residue class mod STRIDE (layout 2: STRIDE = 2 channels interleaved).
// Thread carries f(v) = P v + B over its chunk; returns the value entering the chunk.  smP/smB: 32 cplx each.
// ---------------------------------------------------------------------------------------------
template <bool FWD, int STRIDE>
ION_DEVINL cplx affine_scan_strided_exclusive(cplx P, cplx B, cplx *smP, cplx *smB, int tid, int nthreads, int reach)
{
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5, c = lane % STRIDE;
#pragma unroll
    for (int s = STRIDE; s < 32; s <<= 1) {
        cplx Pp = FWD ? shfl_up_c(P, s) : shfl_down_c(P, s);
        cplx Bp = FWD ? shfl_up_c(B, s) : shfl_down_c(B, s);
        const bool act = FWD ? (lane >= s) : (lane + s < 32);
        if (act) {
            B = c_fma(P, Bp, B);
            P = c_mul(P, Pp);
        }
    }
    cplx win = c_zero();
    if (nw > 1) {
        if (FWD ? (lane >= 32 - STRIDE) : (lane < STRIDE)) {
            smP[warp * STRIDE + c] = P;
            smB[warp * STRIDE + c] = B;
        }
        __syncthreads();
        const int depth = reach > 0 ? reach + 1 : nw;
#pragma unroll 1
        for (int j = depth; j >= 1; --j) {
            const int src = FWD ? warp - j : warp + j;
            if (src >= 0 && src < nw) win = c_fma(smP[src * STRIDE + c], win, smB[src * STRIDE + c]);
        }
    }
    cplx Pe = FWD ? shfl_up_c(P, STRIDE) : shfl_down_c(P, STRIDE);
    cplx Be = FWD ? shfl_up_c(B, STRIDE) : shfl_down_c(B, STRIDE);
    const bool first = FWD ? (lane < STRIDE) : (lane >= 32 - STRIDE);
    return first ? win : c_fma(Pe, win, Be);
}

// layout 1 -> layout 2 for the channel pair (X, Y): the even lane of a lane pair ends up with rows 8p .. 8p+7 of X,
// the odd lane with the same rows of Y
ION_DEVINL void pair_transpose_in(const cplx (&X)[4], const cplx (&Y)[4], cplx (&Z)[8], bool odd)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const cplx give = odd ? X[k] : Y[k];
        const cplx got = make_double2(__shfl_xor_sync(0xffffffffu, give.x, 1), __shfl_xor_sync(0xffffffffu, give.y, 1));
        Z[k] = odd ? got : X[k];
        Z[4 + k] = odd ? Y[k] : got;
    }
}
ION_DEVINL void pair_transpose_out(const cplx (&Z)[8], cplx (&X)[4], cplx (&Y)[4], bool odd)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const cplx give = odd ? Z[k] : Z[4 + k];
        const cplx got = make_double2(__shfl_xor_sync(0xffffffffu, give.x, 1), __shfl_xor_sync(0xffffffffu, give.y, 1));
        X[k] = odd ? got : Z[k];
        Y[k] = odd ? Z[4 + k] : got;
    }
}

// Crank-Nicolson on the 8 consecutive rows of one channel held by this thread (layout 2).
//   wcol: this thread's column of the LU factors in shared memory, row k at wcol[k * T]
//   tocol: tau*off of the thread's rows in shared memory (row k at tocol[k * T/2], the row before the chunk at k = 8)
//   wprev: LU factor of the row before the chunk (0 at the channel start)
//   Pt, Qt: chunk multipliers (forward: e_{-1} e_0 .. e_6, backward: e_0 .. e_7)
ION_DEVINL cplx e_of(double to, cplx w) { return c_make(to * w.y, -to * w.x); }  // -i * to * w

struct NoHook {
    ION_DEVINL void operator()() const {}
};
// `after_first_barrier`: called once every thread of the CTA has passed the forward scan's barrier (the ensemble kernel issues its
// bulk prefetch there: by then all threads have read the staging buffer it overwrites)
// NATURAL: the LU factors sit in shared memory as they sit in global memory -- one channel's four row planes [4][T], moved there by a
// bulk copy -- and wcol points at the thread's column pair: row k of the chunk is plane k & 3, column k >> 2.  Otherwise (cp.async
// staging, permuted on the way in): row k at wcol[k * T].
template <class Hook = NoHook, bool NATURAL = false>
ION_DEVINL void cn8(cplx (&g)[8], const cplx *wcol, int T, const double *tocol, cplx wprev, cplx Pt, cplx Qt, int tid, int nthreads,
                    cplx *sm, int reach, Hook after_first_barrier = Hook())
{
    const int TH = T >> 1;  // tocol[k * TH]: tau*off of row k of the chunk, k = 8: of the row before the chunk
#define to_(k) tocol[(k) * TH]
#define wc_(k) wcol[NATURAL ? (((k) & 3) * T + ((k) >> 2)) : ((k) * T)]
    // forward, zero inflow
    cplx z = g[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) z = c_fma(e_of(to_(k - 1), wc_(k - 1)), z, g[k]);
    const cplx yin = affine_scan_strided_exclusive<true, 2>(Pt, z, sm, sm + 32, tid, nthreads, reach);
    after_first_barrier();
    // forward, true inflow; u = w * y
    cplx u[8];
    cplx wk = wc_(0);
    cplx y = c_fma(e_of(to_(8), wprev), yin, g[0]);
    u[0] = c_mul(wk, y);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const cplx e = e_of(to_(k - 1), wk);
        wk = wc_(k);
        y = c_fma(e, y, g[k]);
        u[k] = c_mul(wk, y);
    }
    // backward, zero inflow
    z = u[7];
#pragma unroll
    for (int k = 6; k >= 0; --k) z = c_fma(e_of(to_(k), wc_(k)), z, u[k]);
    const cplx xin = affine_scan_strided_exclusive<false, 2>(Qt, z, sm + 64, sm + 96, tid, nthreads, reach);
    // backward, true inflow; out = 2 x - g
    cplx x = c_fma(e_of(to_(7), wc_(7)), xin, u[7]);
    g[7] = c_make(fma(2.0, x.x, -g[7].x), fma(2.0, x.y, -g[7].y));
#pragma unroll
    for (int k = 6; k >= 0; --k) {
        x = c_fma(e_of(to_(k), wc_(k)), x, u[k]);
        g[k] = c_make(fma(2.0, x.x, -g[k].x), fma(2.0, x.y, -g[k].y));
    }
#undef to_
#undef wc_
}

// ---------------------------------------------------------------------------------------------
// r-pair rotations on the rows of ONE thread-distributed vector pair (S rotated by +theta, D by -theta):
//   S: [[c, s], [-s, c]] on (i, i+1);  D: [[c, -s], [s, c]]       mesh_operators.py:1247-1408, :384-427
// r_parity 0: pairs (i even, i+1) are thread-local (M even).  r_parity 1: pairs (i odd, i+1); the pair
// (t*M + M-1, (t+1)*M) straddles two threads: both threads compute their half from values exchanged through
// shared memory (xs: 4*T cplx).
// ---------------------------------------------------------------------------------------------
// cos/sin of the r-pair angles of one thread: pairs starting at even local rows k = 0, 2, .. (index k/2), at odd
// local rows k = 1, 3, .., M-1 (index (k-1)/2; the last one is the pair shared with thread t+1) and the pair
// (t*M-1, t*M) shared with thread t-1.
template <int M>
struct RPairAngles {
    double ce[M / 2], se[M / 2], co[M / 2], so[M / 2], cp, sp;
};
template <int M>
ION_DEVINL RPairAngles<M> rpair_angles(const double (&zv)[M], double zprev, double sc)
{
    RPairAngles<M> a;
    double th[M + 1], sn[M + 1], cs[M + 1];
#pragma unroll
    for (int k = 0; k < M; ++k) th[k] = sc * zv[k];
    th[M] = sc * zprev;
    fast_sincos_n<M + 1>(th, sn, cs);
#pragma unroll
    for (int k = 0; k < M; k += 2) a.se[k / 2] = sn[k], a.ce[k / 2] = cs[k];
#pragma unroll
    for (int k = 1; k < M; k += 2) a.so[k / 2] = sn[k], a.co[k / 2] = cs[k];
    a.sp = sn[M];
    a.cp = cs[M];
    return a;
}

template <int M, bool WITH_D>
ION_DEVINL void rpair_layer_even(cplx (&S)[M], cplx (&D)[M], const RPairAngles<M> &ang)
{
#pragma unroll
    for (int k = 0; k < M; k += 2) {
        const double sn = ang.se[k / 2], cs = ang.ce[k / 2];
        cplx s0 = S[k], s1 = S[k + 1];
        S[k] = c_make(fma(cs, s0.x, sn * s1.x), fma(cs, s0.y, sn * s1.y));
        S[k + 1] = c_make(fma(cs, s1.x, -sn * s0.x), fma(cs, s1.y, -sn * s0.y));
        if (WITH_D) {
            cplx d0 = D[k], d1 = D[k + 1];
            D[k] = c_make(fma(cs, d0.x, -sn * d1.x), fma(cs, d0.y, -sn * d1.y));
            D[k + 1] = c_make(fma(cs, d1.x, sn * d0.x), fma(cs, d1.y, sn * d0.y));
        }
    }
}

template <int M, bool WITH_D>
ION_DEVINL void rpair_layer_odd(cplx (&S)[M], cplx (&D)[M], const RPairAngles<M> &ang, int t, int T, cplx *xs)
{
    // publish first and last rows (pre-layer values)
    xs[t] = S[0];
    xs[T + t] = S[M - 1];
    if (WITH_D) {
        xs[2 * T + t] = D[0];
        xs[3 * T + t] = D[M - 1];
    }
    __syncthreads();
    cplx nS0 = (t + 1 < T) ? xs[t + 1] : c_zero();         // next thread's first row
    cplx pSL = (t > 0) ? xs[T + t - 1] : c_zero();         // previous thread's last row
    cplx nD0 = c_zero(), pDL = c_zero();
    if (WITH_D) {
        nD0 = (t + 1 < T) ? xs[2 * T + t + 1] : c_zero();
        pDL = (t > 0) ? xs[3 * T + t - 1] : c_zero();
    }
    // no second barrier: every caller runs a CTA-wide barrier (the scans of a Crank-Nicolson solve) or ends the kernel before xs is
    // written again -- one odd layer per h2_pair, and two h2_pair / line sweeps are always separated by a solve
    // interior odd pairs (k, k+1), k = 1, 3, ..., M-3
#pragma unroll
    for (int k = 1; k + 1 < M; k += 2) {
        const double sn = ang.so[k / 2], cs = ang.co[k / 2];
        cplx s0 = S[k], s1 = S[k + 1];
        S[k] = c_make(fma(cs, s0.x, sn * s1.x), fma(cs, s0.y, sn * s1.y));
        S[k + 1] = c_make(fma(cs, s1.x, -sn * s0.x), fma(cs, s1.y, -sn * s0.y));
        if (WITH_D) {
            cplx d0 = D[k], d1 = D[k + 1];
            D[k] = c_make(fma(cs, d0.x, -sn * d1.x), fma(cs, d0.y, -sn * d1.y));
            D[k + 1] = c_make(fma(cs, d1.x, sn * d0.x), fma(cs, d1.y, sn * d0.y));
        }
    }
    {   // my last row is the LOWER member of (t*M+M-1, (t+1)*M); its angle is 0 if there is no such pair
        const double sn = ang.so[M / 2 - 1], cs = ang.co[M / 2 - 1];
        cplx s0 = S[M - 1];
        S[M - 1] = c_make(fma(cs, s0.x, sn * nS0.x), fma(cs, s0.y, sn * nS0.y));
        if (WITH_D) {
            cplx d0 = D[M - 1];
            D[M - 1] = c_make(fma(cs, d0.x, -sn * nD0.x), fma(cs, d0.y, -sn * nD0.y));
        }
    }
    {   // my first row is the UPPER member of (t*M-1, t*M); angle zprev (0 for t = 0)
        const double sn = ang.sp, cs = ang.cp;
        cplx s1 = S[0];
        S[0] = c_make(fma(cs, s1.x, -sn * pSL.x), fma(cs, s1.y, -sn * pSL.y));
        if (WITH_D) {
            cplx d1 = D[0];
            D[0] = c_make(fma(cs, d1.x, sn * pDL.x), fma(cs, d1.y, sn * pDL.y));
        }
    }
}

// Hadamard over the l-pair without its 1/sqrt(2) on the way in (the bricks are linear) ...
template <int M>
ION_DEVINL void hadamard_in(cplx (&A)[M], cplx (&B)[M])
{
#pragma unroll
    for (int k = 0; k < M; ++k) {
        cplx a = A[k], b = B[k];
        A[k] = c_add(a, b);  // sqrt(2) S
        B[k] = c_sub(a, b);  // sqrt(2) D
    }
}
// ... and both factors (1/2, exact) on the way out: 3 instructions per component pair
template <int M>
ION_DEVINL void hadamard_out(cplx (&A)[M], cplx (&B)[M])
{
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const cplx h = c_scale(A[k], 0.5), d = B[k];
        A[k] = c_make(fma(0.5, d.x, h.x), fma(0.5, d.y, h.y));
        B[k] = c_make(fma(-0.5, d.x, h.x), fma(-0.5, d.y, h.y));
    }
}

// Hadamard over the l-pair, the two r-sublayers, Hadamard back  (SimilarityOperator, mesh_operators.py:150-204)
template <int M>
ION_DEVINL void h2_pair(cplx (&A)[M], cplx (&B)[M], const RPairAngles<M> &ang, bool reverse, int t, int T, cplx *xs)
{
    hadamard_in<M>(A, B);
    if (!reverse) {
        rpair_layer_even<M, true>(A, B, ang);
        rpair_layer_odd<M, true>(A, B, ang, t, T, xs);
    } else {
        rpair_layer_odd<M, true>(A, B, ang, t, T, xs);
        rpair_layer_even<M, true>(A, B, ang);
    }
    hadamard_out<M>(A, B);
}

// =============================================================================================
// Unit kernels.  grid = (units, batch), block = T.
// =============================================================================================
enum : int {
    PROG_ROT = 0,        // rotation by (s_a + s_b) [+ mask]                       -- LEN even sweep, VEL h1 sweeps
    PROG_ROT_CN_ROT = 1, // rotation(s_a), CN on both channels, rotation(s_a)      -- LEN odd sweep around CN
    PROG_H2 = 2,         // Hadamard r-pair bricks                                  -- VEL h2 on even l-pairs
    PROG_H2_CN_H2 = 3,   // bricks, CN, bricks reversed                             -- VEL h2 on odd l-pairs around CN
    PROG_CN = 4,         // CN on every channel (unit = channel)                    -- generic path
    PROG_LINE_SO_LEN = 5,// exp(-i s w_z) * CN * exp(-i s w_z) [+ mask]             -- LineMesh SO length gauge
    PROG_LINE_SO_VEL = 6,// r-pair rotations even, odd, CN, odd, even [+ mask]      -- LineMesh SO velocity gauge
    PROG_LINE_CN = 7,    // CN with H = H0 + diag(s w_z): pivots rebuilt every step     -- LineMesh CN (ADI) length gauge
    PROG_LEN_STEP = 8,   // even rotation(s_a + s_b) [+ mask] of the unit's channels with their READ-ONLY even-pair partners,
                         // then rotation(s_a), CN, rotation(s_a) on the odd pair; out of place -- one pass per LEN step
    PROG_LEN_STEP_OBS = 9,  // the same with the observation of the PREVIOUS step fused in: even rotation(s_b), mask, reductions over
                            // the unit's own channels, even rotation(s_a), ... (north_star 4; see obs_channel)
    PROG_LEN_STEP_HALO = 10,  // PROG_LEN_STEP of an l-block shard cut at odd channels with the halo exchange FUSED in: the CTAs of the
                              // block's first / last pair store their boundary channel into the neighbour's memory over NVLink
                              // (epilogue) and read their ghost partner from the slot the neighbour filled one step earlier (prologue)
};

// One member of an l-pair rotation [[c, -i s], [-i s, c]] (the matrix is symmetric: both members use the same formula)
template <int M>
ION_DEVINL void rotate_member(cplx (&X)[M], const cplx (&partner)[M], const RotAngles<M> &ang)
{
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const cplx x = X[k], q = partner[k];
        X[k] = c_make(fma(ang.c[k], x.x, ang.s[k] * q.y), fma(ang.c[k], x.y, -ang.s[k] * q.x));
    }
}
// even-pair partner of local channel c: (index, coefficient index) -- channel c with GLOBAL index gl pairs with gl^1
ION_DEVINL bool even_partner(const UnitParams &p, int c, int &pc, int &ci)
{
    const int gl = p.l_begin + c;
    pc = (gl & 1) ? c - 1 : c + 1;
    ci = (gl & 1) ? gl - 1 : gl;
    return pc >= 0 && pc < p.L;
}

// both members of an even pair [[c, -i s], [-i s, c]] from their old values (the partner is read-only in memory, but the
// fused observation needs ITS post-rotation value too: the next half-rotation of X acts on the rotated pair)
template <int M>
ION_DEVINL void rotate_both(cplx (&X)[M], cplx (&Q)[M], const RotAngles<M> &ang)
{
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const cplx x = X[k], q = Q[k];
        X[k] = c_make(fma(ang.c[k], x.x, ang.s[k] * q.y), fma(ang.c[k], x.y, -ang.s[k] * q.x));
        Q[k] = c_make(fma(ang.c[k], q.x, ang.s[k] * x.y), fma(ang.c[k], q.y, -ang.s[k] * x.x));
    }
}

// Fused observation of ONE channel held by the CTA (rows k*T + t of thread t in X): norm, <r>, norm within radii ->
// obs_partial[b][l] (k_observe's layout; the <z> and <H0> slots are zero: those observables take the unfused path), inner
// products with the channel's test states -> obs_ip.  CTA-wide reductions in a fixed order (deterministic).  sm: >= 32 * 10 doubles.
#ifndef ION_MAX_RADII
#define ION_MAX_RADII 8
#endif
template <int M>
ION_DEVINL void obs_channel(const UnitParams &p, const cplx (&X)[M], int l, int b, int t, bool ok, int tl, int Tc, double *sm)
{
    const int T = p.T;
    double n2[M], rr[M], acc[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < M; ++k) {
        n2[k] = ok ? c_abs2(X[k]) : 0.0;  // padding rows are zero
        rr[k] = (ok && p.obs_rvec) ? p.obs_rvec[k * T + t] : 0.0;
        acc[0] += n2[k];
        acc[1] += rr[k] * n2[k];
    }
    block_sum<2>(acc, sm, tl, Tc);
    double *out = p.obs_partial + ((size_t)b * p.L + l) * (4 + p.obs_n_radii);
    if (tl == 0) {
        out[0] = acc[0];
        out[1] = acc[1];
        out[2] = 0.0;
        out[3] = 0.0;
    }
    for (int q = 0; q < p.obs_n_radii; ++q) {  // norm within radius q (mesh/data.py:419-422)
        double w[1] = {0.0};
#pragma unroll
        for (int k = 0; k < M; ++k) w[0] += (rr[k] <= p.obs_radii[q]) ? n2[k] : 0.0;
        block_sum<1>(w, sm, tl, Tc);
        if (tl == 0) out[4 + q] = w[0];
    }
    if ((p.obs_what & 2u) && p.obs_n_states > 0) {
        for (int si = p.obs_state_first[l]; si < p.obs_state_first[l + 1]; ++si) {  // uniform over the CTA
            const int st = p.obs_state_order[si];
            const cplx *row = p.obs_state_rows + (size_t)st * M * T;
            double ip[2] = {0.0, 0.0};
#pragma unroll
            for (int k = 0; k < M; ++k) {
                if (ok) {
                    const cplx a = row[k * T + t], x = X[k];
                    ip[0] += a.x * x.x + a.y * x.y;
                    ip[1] += a.x * x.y - a.y * x.x;
                }
            }
            block_sum<2>(ip, sm, tl, Tc);
            if (tl == 0) {
                p.obs_ip[((size_t)b * p.obs_n_states + st) * 2 + 0] = ip[0] * p.obs_ipm;
                p.obs_ip[((size_t)b * p.obs_n_states + st) * 2 + 1] = ip[1] * p.obs_ipm;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LU factors of (1 + i tau (H0 + diag(E w_z))) built on the fly (LineMesh Crank-Nicolson in the length gauge:
// evolution_methods.py:49-77 with mesh_operators.py:271-298, :320-327 -- the matrix changes every step and differs
// between the members of an ensemble, so nothing can be precomputed).  The pivot recurrence is a Moebius map per row;
// each thread composes the maps of its M rows, a Kogge-Stone scan of 2x2 matrices over the warp gives every lane the
// pivot entering its chunk, and the thread then rebuilds its M pivots.  Across warps only the neighbour's last pivot
// is needed: a pivot forgets its starting value at the rate |o/p|^2 per row (< 1e-30 over a warp; the host checks the
// same decay bound as for r-segments before it accepts this program).
// ---------------------------------------------------------------------------------------------
template <int M>
ION_DEVINL void line_cn_factors(CnFactors<M> &f, const cplx (&D)[M], const double (&toff)[M], double toff_prev, int tl, cplx *sm)
{
    const int lane = tl & 31, warp = tl >> 5;
    double o2[M];
    o2[0] = toff_prev * toff_prev;
#pragma unroll
    for (int k = 1; k < M; ++k) o2[k] = toff[k - 1] * toff[k - 1];
    Mat2 X;
    X.a = D[0];
    X.b = c_make(o2[0], 0.0);
    X.c = c_make(1.0, 0.0);
    X.d = c_zero();
#pragma unroll
    for (int k = 1; k < M; ++k) {
        Mat2 Y;
        Y.a = c_make(fma(o2[k], X.c.x, fma(D[k].x, X.a.x, -D[k].y * X.a.y)), fma(o2[k], X.c.y, fma(D[k].x, X.a.y, D[k].y * X.a.x)));
        Y.b = c_make(fma(o2[k], X.d.x, fma(D[k].x, X.b.x, -D[k].y * X.b.y)), fma(o2[k], X.d.y, fma(D[k].x, X.b.y, D[k].y * X.b.x)));
        Y.c = X.a;
        Y.d = X.b;
        X = Y;
    }
    mat_normalize(X);
    // five steps: every lane ends up with the composition over all the threads of its warp before it (up to 128 rows).  A pivot
    // forgets its starting value at the rate |o / p|^2 per row.  Field-free that is < 1e-30 over 64 rows under the decay bound this
    // program requires -- but the diagonal also carries tau E w_z, and where the field's potential cancels the kinetic diagonal
    // the rate rises to ~0.7 per row (config 2 at 10 J/cm^2: 64 rows left 1e-10, measured against the oracle); 128 rows: 1e-20.
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        Mat2 Y = mat_shfl_up(X, s);
        if (lane >= s) {
            X = mat_mul(X, Y);
            mat_normalize(X);
        }
    }
    // last pivot of this warp when nothing precedes it (start value "infinity" = (1, 0)): n/d = a/c
    if (lane == 31) sm[warp] = c_div(X.a, X.c);
    __syncthreads();
    const bool inf_in = (warp == 0);
    const cplx pin_warp = inf_in ? c_zero() : sm[warp - 1];
    Mat2 Y = mat_shfl_up(X, 1);
    cplx wprev;  // 1 / (pivot of the row before this thread's first row); 0 when there is none
    if (lane == 0) {
        wprev = inf_in ? c_zero() : c_inv(pin_warp);
    } else {
        cplx n = inf_in ? Y.a : c_fma(Y.a, pin_warp, Y.b);
        cplx d = inf_in ? Y.c : c_fma(Y.c, pin_warp, Y.d);
        wprev = c_div(d, n);
    }
    f.wprev = wprev;
    cplx wp = wprev;
#pragma unroll
    for (int k = 0; k < M; ++k) {
        cplx piv = c_make(fma(o2[k], wp.x, D[k].x), fma(o2[k], wp.y, D[k].y));
        wp = c_inv(piv);
        f.w[k] = wp;
    }
    // chunk aggregates of the affine recurrences (k_aggregates, on the fly)
    cplx P = c_make(toff_prev * wprev.y, -toff_prev * wprev.x), Q = c_make(1.0, 0.0);
#pragma unroll
    for (int k = 0; k < M; ++k) {
        cplx e = c_make(toff[k] * f.w[k].y, -toff[k] * f.w[k].x);
        if (k < M - 1) P = c_mul(P, e);
        Q = c_mul(Q, e);
    }
    f.P = P;
    f.Q = Q;
}

// register budget: the programs without a Crank-Nicolson solve are asked to fit two CTAs of TMAX threads per SM
// (<= 64 registers at TMAX = 512) so that one CTA's loads overlap the other's arithmetic
//
// r-SEGMENTS.  A channel longer than one CTA can hold (r_points > 4096) is cut into S segments of T_seg threads.
// Each CTA also computes H halo threads (H*M rows) on either side: the r-pair bricks need their neighbours, and the
// Crank-Nicolson recurrences are simply started from zero at the edge of the halo.  The halo rows of one CTA are the
// interior rows of another, so every segmented kernel that reads a halo runs OUT OF PLACE (psi -> psi_out, the engine
// swaps the buffers): in place, a CTA of a later wave would read rows its neighbour has already advanced.  That is exact to < 1e-30
// because the LU multipliers decay geometrically -- the host verifies (k_scan_bound) that their product over any 32
// threads is below 1e-30 before it allows S > 1.  Halo results are discarded; only interior threads store.
// The length-gauge programs (no r-pair bricks) take HALF-WARP halos (16 threads = 64 rows) when the product over any aligned
// 16 threads is below 1e-18 -- two orders below the rounding of the values it multiplies: 224 interior threads of 256.
#ifndef ION_LU_BULK
#define ION_LU_BULK 1
#endif
template <int M, int PROG, int TMAX, bool SEG>
#ifndef ION_PAIR_MINB
#define ION_PAIR_MINB 1
#endif
// (eight rows per thread in r-segments -- long LineMesh channels, see engine.cu -- are capped at 128 registers: two 256-thread CTAs per SM)
__global__ void __launch_bounds__(TMAX, ((PROG == PROG_ROT || PROG == PROG_H2) && TMAX <= 512) ? (M <= 4 ? 1024 / TMAX : 2)
                                                                                            : ((M <= 4 && TMAX <= 256) ? 2 : (TMAX == 512 && M == 4 ? ION_PAIR_MINB : ((M == 8 && SEG) ? 2 : 1))))
    k_unit(const UnitParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm_scan = reinterpret_cast<cplx *>(smem_raw);  // 2 channels x 128 cplx
    cplx *xs = sm_scan + 256;                            // 4*blockDim cplx (r-pair exchange)

    const int tl = threadIdx.x, Tc = blockDim.x;         // index inside the CTA: scans and exchanges
    const int T = p.T;                                   // row stride of the layout
    // SEG == false (one CTA per channel, the common case): all of this folds away at compile time
    const int seg = SEG ? (int)(blockIdx.x % p.S) : 0;
    int ux = SEG ? (int)(blockIdx.x / p.S) : (int)blockIdx.x;
    // boundary units first: what they send is under way long before the neighbour's next launch asks for it
    if (PROG == PROG_LEN_STEP_HALO && p.boundary_first) ux = ux == 0 ? 0 : (ux == 1 ? p.n_units - 1 : ux - 1);
    const int unit = p.unit0 + ux * p.unit_stride;
    const int t = SEG ? seg * p.T_seg - p.H + tl : tl;   // thread index inside the channel: addressing
    const bool ok = SEG ? ((t >= 0) && (t < T)) : true;
    const bool mine = SEG ? (ok && (tl >= p.H) && (tl < p.H + p.T_seg)) : true;
    const int b = blockIdx.y;
    int l0;
    bool pair = true;
    if (PROG == PROG_CN || PROG == PROG_LINE_SO_LEN || PROG == PROG_LINE_SO_VEL || PROG == PROG_LINE_CN) {
        l0 = unit;
        pair = false;
    } else {
        unit_channels(p, unit, l0, pair);
    }
    const size_t chan = (size_t)M * T;
    cplx *base = p.psi + ((size_t)b * p.L + l0) * chan;
    const double sa = p.scal_a ? p.scal_a[b] : 0.0;
    const double sb = p.scal_b ? p.scal_b[b] : 0.0;

    cplx A[M], B[M];
    pdl_launch_dependents();  // see common.cuh: everything up to pdl_wait() is independent of psi

    if (PROG == PROG_ROT) {
        if (!pair && !(p.flags & F_MASK)) return;
        if (!mine) return;  // point-wise in r: no halo needed (launched with H = 0)
        RotAngles<M> ang;
        double mk[M];
        if (pair) {
            double vec[M];
            load_vec<M>(vec, p.vec, T, t, true);
            ang = rot_angles<M>(vec, (sa + sb) * p.cl[p.l_begin + l0]);
        }
        if (p.flags & F_MASK) load_vec<M>(mk, p.mask, T, t, true);
        pdl_wait();
        load_rows<M>(A, base, T, t, true);
        if (pair) {
            load_rows<M>(B, base + chan, T, t, true);
            if (p.flags & F_REAL_ROT) rotate_pair<M, true>(A, B, ang);
            else rotate_pair<M, false>(A, B, ang);
        }
        if (p.flags & F_MASK) {
#pragma unroll
            for (int k = 0; k < M; ++k) {
                A[k] = c_scale(A[k], mk[k]);
                if (pair) B[k] = c_scale(B[k], mk[k]);
            }
        }
        store_rows<M>(A, base, T, t, true);
        if (pair) store_rows<M>(B, base + chan, T, t, true);
        return;
    }

    if (PROG == PROG_H2) {
        cplx *obase = p.psi_out + ((size_t)b * p.L + l0) * chan;
        if (!pair) {  // nothing to do for an unpaired channel; out of place it still has to reach the other buffer
            if (p.psi_out != p.psi) {
                pdl_wait();
                load_rows<M>(A, base, T, t, mine);
                store_rows<M>(A, obase, T, t, mine);
            }
            return;
        }
        double zv[M];
        load_vec<M>(zv, p.zvec, T, t, ok);
        double sc = sa * p.cl2[p.l_begin + l0];
        const RPairAngles<M> ang = rpair_angles<M>(zv, ok ? p.zprev[t] : 0.0, sc);
        pdl_wait();
        load_rows<M>(A, base, T, t, ok);
        load_rows<M>(B, base + chan, T, t, ok);
        h2_pair<M>(A, B, ang, (p.flags & F_H2_REVERSE) != 0, tl, Tc, xs);
        store_rows<M>(A, obase, T, t, mine);
        store_rows<M>(B, obase + chan, T, t, mine);
        return;
    }

    // ---- programs containing Crank-Nicolson ----
    // pairs: both channels are solved together in layout 2 (see above); single channels and r-segments keep layout 1
    constexpr bool LENSTEP = (PROG == PROG_LEN_STEP || PROG == PROG_LEN_STEP_OBS || PROG == PROG_LEN_STEP_HALO), OBS = (PROG == PROG_LEN_STEP_OBS);
    constexpr bool HALO = (PROG == PROG_LEN_STEP_HALO);
    // (r-segments included: the thread's position in the channel is t = segment offset + tl, halo threads beyond either end of the
    // channel -- !ok -- carry identity factors and zero multipliers, and the recurrences start from zero at the edge of the halo)
    constexpr bool L2CN = (M == 4) && (TMAX <= 512) && (PROG == PROG_ROT_CN_ROT || PROG == PROG_H2_CN_H2 || LENSTEP);
#ifdef ION_EXP_CLOCKS  // timing-only instrumentation: phase time stamps of two CTAs (first and second wave)
    long long ck[8];
#define ION_CK(i) ck[i] = clock64()
#else
#define ION_CK(i)
#endif
    ION_CK(0);
    cplx *obase = p.psi_out + ((size_t)b * p.L + l0) * chan;
    if constexpr (L2CN) if (pair) {
        // LU factors of both channels.  One 512-thread CTA per pair: each channel's four row planes are one contiguous block in global memory,
        // moved by the TMA engine (two cp.async.bulk issued by thread 0 at the very top, completion on an mbarrier) into [4][Tc] per
        // channel; the upper channel is shifted by one element so that the two lanes of a lane pair -- even lane: lower channel, odd lane:
        // upper channel, same columns -- read different banks.  r-segments: cp.async per thread, permuted to row k at wsm[k * Tc + tl].
        constexpr bool LUB = !SEG && TMAX == 512 && (ION_LU_BULK != 0);  // (smaller CTAs: no gain, 500 x 50 measured 3 % slower)
        cplx *wsm = xs + 4 * Tc;                                      // [8][Tc] (+ 1)
        double *tosm = reinterpret_cast<double *>(wsm + 8 * Tc + 1);  // [9][Tc/2] tau*off of the chunk's rows (+ the row before it)
        unsigned long long *lubar = reinterpret_cast<unsigned long long *>(tosm + 9 * (Tc >> 1) + ((9 * (Tc >> 1)) & 1));
        if constexpr (LUB) {
            if (tl == 0) {
                mbar_init(lubar, 1);
                mbar_expect_tx(lubar, (unsigned)(2 * chan * sizeof(cplx)));
                bulk_g2s(wsm, p.w + (size_t)l0 * chan, (unsigned)(chan * sizeof(cplx)), lubar);
                bulk_g2s(wsm + chan + 1, p.w + (size_t)(l0 + 1) * chan, (unsigned)(chan * sizeof(cplx)), lubar);
            }
            __syncthreads();  // the barrier is initialised before anybody waits on it
        }
        const int pp = tl >> 1, TH = Tc >> 1;
        const bool odd = (tl & 1) != 0;
        // The small coefficient loads go FIRST: the 64 KB of LU factors of a 512-thread CTA keep the SM's load path busy for
        // ~1500 cycles, and everything queued behind them (and the trigonometry that waits for it) would be exposed in every CTA
        // of a second wave, where no previous kernel's tail hides the prologue.
        const int tp = t - (tl & 1);  // position in the channel of the lane pair's first thread (tp == 2 pp without r-segments)
        const bool ok0 = SEG ? (tp >= 0 && tp < T) : true, ok1 = SEG ? (tp + 1 >= 0 && tp + 1 < T) : true;
        double tov[9];
        if (!odd) {
#pragma unroll
            for (int k = 0; k < 8; ++k) tov[k] = ((k >> 2) ? ok1 : ok0) ? p.toff[(k & 3) * T + tp + (k >> 2)] : 0.0;
            tov[8] = ok0 ? p.toff_prev[tp] : 0.0;
        }
        cplx P8, Q8;  // multipliers of the 8-row chunk = product of the two 4-row ones
        cplx aP0, aP1, aQ0, aQ1;
        {
            const size_t ch = (size_t)(l0 + (odd ? 1 : 0)) * T + tp;
            aP0 = ok0 ? ld_c(p.aggP + ch) : c_zero(), aP1 = ok1 ? ld_c(p.aggP + ch + 1) : c_zero();
            aQ0 = ok0 ? ld_c(p.aggQ + ch) : c_zero(), aQ1 = ok1 ? ld_c(p.aggQ + ch + 1) : c_zero();
        }
        // LU factor of the row before the chunk (the previous lane pair's last row, possibly another warp's; at the first thread of a
        // segment its inflow is zero by construction, so the factor does not matter there)
        const cplx wprev = (pp > 0 && tp > 0 && tp - 1 < T) ? ld_c(p.w + (size_t)(l0 + (odd ? 1 : 0)) * chan + 3 * (size_t)T + tp - 1) : c_zero();
        double cvec[M], czp = 0.0, kap0, kapA = 0.0, kapB = 0.0;
        if (PROG == PROG_ROT_CN_ROT || LENSTEP) {
            load_vec<M>(cvec, p.vec, T, t, ok);
            kap0 = sa * p.cl[p.l_begin + l0];
            if (LENSTEP) {  // a pair always has both even-pair partners: l0 - 1 and l0 + 2
                // OBS: the previous step's half (s_b) first, on its own -- the observed state sits between the two halves
                kapA = (OBS ? sb : sa + sb) * p.cl[p.l_begin + l0 - 1];
                kapB = (OBS ? sb : sa + sb) * p.cl[p.l_begin + l0 + 1];
            }
        } else {
            load_vec<M>(cvec, p.zvec, T, t, ok);
            czp = ok ? p.zprev[t] : 0.0;
            kap0 = sa * p.cl2[p.l_begin + l0];
        }
        if constexpr (!LUB) {   // coalesced reads of both channels' factors, permuted on the way into shared memory
            const cplx *w0 = p.w + (size_t)l0 * chan + t;
            cplx *dst = wsm + (size_t)(4 * (tl & 1)) * Tc + (tl & ~1);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                if (ok) {
                    cp_async16(dst + (size_t)k4 * Tc, w0 + (size_t)k4 * T);
                    cp_async16(dst + (size_t)k4 * Tc + 1, w0 + chan + (size_t)k4 * T);
                } else {  // beyond the channel: identity factors (read by this thread and its lane-pair partner only, after the __syncwarp below)
                    dst[(size_t)k4 * Tc] = c_make(1.0, 0.0);
                    dst[(size_t)k4 * Tc + 1] = c_make(1.0, 0.0);
                }
            }
            cp_async_commit();
        }
        if (!odd) {
#pragma unroll
            for (int k = 0; k < 9; ++k) tosm[k * TH + pp] = tov[k];
        }
        P8 = c_mul(aP0, aP1);
        Q8 = c_mul(aQ0, aQ1);
        RotAngles<M> rang, eangA, eangB;
        RPairAngles<M> pang;
        if (PROG == PROG_ROT_CN_ROT || LENSTEP) {
            rang = rot_angles_auto<M>(cvec, p.vec_dv, kap0);
            if (LENSTEP) {
                eangA = rot_angles_auto<M>(cvec, p.vec_dv, kapA);
                eangB = rot_angles_auto<M>(cvec, p.vec_dv, kapB);
            }
        }
        // PROG_H2_CN_H2 follows k_slab, whose one CTA per SM holds the whole register file until it exits: no CTA of this kernel is
        // resident early, so nothing placed before the wait overlaps the previous kernel.  Its trigonometry is therefore issued
        // AFTER the psi loads, where it hides their latency (ION_H2_TRIG_FIRST restores the old order for A/B timing).
#ifdef ION_H2_TRIG_FIRST
        if (PROG == PROG_H2_CN_H2) pang = rpair_angles<M>(cvec, czp, kap0);
#endif
        ION_CK(1);
        pdl_wait();
        ION_CK(2);
        bool blo = false, bhi = false;    // HALO: this CTA's pair reads the lower / upper ghost channel and sends the boundary channel
        unsigned long long ksent = 0ull;  // HALO: launches so far
        load_rows<M>(A, base, T, t, ok);
        load_rows<M>(B, base + chan, T, t, ok);
#ifndef ION_H2_TRIG_FIRST
        if (PROG == PROG_H2_CN_H2) pang = rpair_angles<M>(cvec, czp, kap0);
#endif
        if constexpr (OBS) {
            // psi_n = mask E_e(s_b) (...) on the even pairs (l0 - 1, l0) and (l0 + 1, l0 + 2): both members, because this step's
            // own half-rotation E_e(s_a) then acts on the rotated, masked pair.  The unit's own channels l0, l0 + 1 are observed.
            double mk[M];
#pragma unroll
            for (int k = 0; k < M; ++k) mk[k] = 1.0;
            if (p.flags & F_MASK) load_vec<M>(mk, p.mask, T, t, ok);
            double *osm = reinterpret_cast<double *>(xs);
            cplx Q[M];
            load_rows<M>(Q, base - chan, T, t, ok);
            rotate_both<M>(A, Q, eangA);
#pragma unroll
            for (int k = 0; k < M; ++k) A[k] = c_scale(A[k], mk[k]), Q[k] = c_scale(Q[k], mk[k]);
            obs_channel<M>(p, A, l0, b, t, mine, tl, Tc, osm);
            rotate_member<M>(A, Q, rot_angles_auto<M>(cvec, p.vec_dv, sa * p.cl[p.l_begin + l0 - 1]));
            load_rows<M>(Q, base + 2 * chan, T, t, ok);
            rotate_both<M>(B, Q, eangB);
#pragma unroll
            for (int k = 0; k < M; ++k) B[k] = c_scale(B[k], mk[k]), Q[k] = c_scale(Q[k], mk[k]);
            obs_channel<M>(p, B, l0 + 1, b, t, mine, tl, Tc, osm);
            rotate_member<M>(B, Q, rot_angles_auto<M>(cvec, p.vec_dv, sa * p.cl[p.l_begin + l0 + 1]));
        } else if (LENSTEP) {
            cplx Q[M];
            const cplx *qlo = base - chan, *qhi = base + 2 * chan;
            if constexpr (HALO) {
                // FUSED HALO EXCHANGE, receiving side.  k = launches of this program so far (every launch sends, so the neighbour's
                // arrival counter after its launch k-1 is S * k); the ghost partner is in the slot (k-1) & 1 the neighbour's launch
                // k-1 stored into.  No hand-shake for the slot's reuse: the neighbour's launch k+1 overwrites it only after it has seen
                // the S arrivals of MY launch k, which every CTA of mine sends after its reads.
                blo = (unit == p.hf_unit[0]), bhi = (unit == p.hf_unit[1]);
                if (blo || bhi) {
                    ksent = p.hf_sent[(blo ? 0 : 1) * p.S + seg];
                    if (p.hf_consume) {
                        __shared__ int hf_ok;
                        if (tl == 0) {
                            const unsigned long long need = (unsigned long long)p.S * ksent;
                            bool good = true;
                            if (blo) good = halo_spin(p.hf_flags + HF_FARRIVE + 0, need, p.hf_flags, p.hf_spin_limit);
                            if (bhi && good) good = halo_spin(p.hf_flags + HF_FARRIVE + 1, need, p.hf_flags, p.hf_spin_limit);
                            if (!good)  // tell both neighbours: they must not keep stepping with a ghost channel that never arrived here
                                for (int q = 0; q < 2; ++q)
                                    if (p.hf_peer_flags[q]) st_release_sys(p.hf_peer_flags[q] + HF_ABORT, 1ull);
                            hf_ok = good ? 1 : 0;
                        }
                        __syncthreads();
                        const size_t slot = (size_t)((ksent + 1ull) & 1ull) * chan;
                        if (blo) qlo = p.hf_my_fstage[0] + slot;
                        if (bhi) qhi = p.hf_my_fstage[1] + slot;
                    }
                }
            }
            if (HALO && blo && p.hf_consume) load_rows_cg<M>(Q, qlo, T, t, ok);
            else load_rows<M>(Q, qlo, T, t, ok);
            rotate_member<M>(A, Q, eangA);
            if (HALO && bhi && p.hf_consume) load_rows_cg<M>(Q, qhi, T, t, ok);
            else load_rows<M>(Q, qhi, T, t, ok);
            rotate_member<M>(B, Q, eangB);
            if (p.flags & F_MASK) {
                double mk[M];
                load_vec<M>(mk, p.mask, T, t, ok);
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    A[k] = c_scale(A[k], mk[k]);
                    B[k] = c_scale(B[k], mk[k]);
                }
            }
        }
        if (PROG == PROG_ROT_CN_ROT || LENSTEP) rotate_pair<M, false>(A, B, rang);
        else h2_pair<M>(A, B, pang, false, tl, Tc, xs);  // (oe, oo)
        ION_CK(3);
        if constexpr (LUB) mbar_wait(lubar, 0u);
        else cp_async_wait_all();
        __syncwarp();  // a thread reads factors (tau*off) staged by itself and by its lane-pair partner only (wprev comes from global memory)
        ION_CK(4);
        {
            cplx Z[8];
            pair_transpose_in(A, B, Z, odd);
            if constexpr (LUB) cn8<NoHook, true>(Z, wsm + (odd ? chan + 1 : 0) + 2 * pp, Tc, tosm + pp, wprev, P8, Q8, tl, Tc, sm_scan, p.short_scan);
            else cn8(Z, wsm + tl, Tc, tosm + pp, wprev, P8, Q8, tl, Tc, sm_scan, p.short_scan);
            pair_transpose_out(Z, A, B, odd);
        }
        ION_CK(5);
        if (PROG == PROG_ROT_CN_ROT || LENSTEP) rotate_pair<M, false>(A, B, rang);
        else h2_pair<M>(A, B, pang, true, tl, Tc, xs);  // (oo, oe)
        ION_CK(6);
        store_rows<M>(A, obase, T, t, mine);
        store_rows<M>(B, obase + chan, T, t, mine);
        if constexpr (HALO) {
            // FUSED HALO EXCHANGE, sending side: the block's first channel goes to the lower neighbour, its last channel to the upper
            // one -- straight from the registers that hold the result, into the neighbour's slot k & 1, then ONE arrival per CTA
            if (blo || bhi) {
                const size_t slot = (size_t)(ksent & 1ull) * chan;
                if (blo && p.hf_peer_fstage[0]) store_rows<M>(A, p.hf_peer_fstage[0] + slot, T, t, mine);
                if (bhi && p.hf_peer_fstage[1]) store_rows<M>(B, p.hf_peer_fstage[1] + slot, T, t, mine);
                __threadfence_system();
                __syncthreads();
                if (tl == 0) {
                    if (blo) {
                        if (p.hf_peer_flags[0]) red_release_sys_add(p.hf_peer_flags[0] + HF_FARRIVE + 1, 1ull);
                        p.hf_sent[0 * p.S + seg] = ksent + 1ull;
                    }
                    if (bhi) {
                        if (p.hf_peer_flags[1]) red_release_sys_add(p.hf_peer_flags[1] + HF_FARRIVE + 0, 1ull);
                        p.hf_sent[1 * p.S + seg] = ksent + 1ull;
                    }
                }
            }
        }
        ION_CK(7);
#ifdef ION_EXP_CLOCKS
        if ((tl == 0 || tl == 288) && (blockIdx.x == 3 || blockIdx.x == 200) && blockIdx.y == 0)
            printf("CK prog %d cta %d t %d: prologue %lld pdlwait %lld load+op1 %lld factorwait %lld cn %lld op2 %lld store %lld total %lld\n", PROG, (int)blockIdx.x, tl,
                   ck[1] - ck[0], ck[2] - ck[1], ck[3] - ck[2], ck[4] - ck[3], ck[5] - ck[4], ck[6] - ck[5], ck[7] - ck[6], ck[7] - ck[0]);
#endif
        return;
    }
    double toff[M];
    load_vec<M>(toff, p.toff, T, t, ok);
    const double toff_prev = ok ? p.toff_prev[t] : 0.0;
    CnFactors<M> fA, fB;
    if (PROG != PROG_LINE_CN) {
        const bool single_channel_prog = (PROG == PROG_LINE_SO_LEN || PROG == PROG_LINE_SO_VEL);
        const size_t lw = single_channel_prog ? 0 : (size_t)l0;
        cn_load<M>(fA, p.w + lw * chan, p.aggP + lw * T, p.aggQ + lw * T, t, T, ok);
        if (pair) cn_load<M>(fB, p.w + (lw + 1) * chan, p.aggP + (lw + 1) * T, p.aggQ + (lw + 1) * T, t, T, ok);
    }
    const int short_scan = p.short_scan;
    // trigonometry of the programs that rotate before the solve: also independent of psi
    RotAngles<M> rang;
    RPairAngles<M> pang;
    if ((PROG == PROG_ROT_CN_ROT || LENSTEP) && pair) {
        double vec[M];
        load_vec<M>(vec, p.vec, T, t, ok);
        rang = rot_angles<M>(vec, sa * p.cl[p.l_begin + l0]);
    }
    // PROG_LEN_STEP: even half-rotations of A (and B) with their read-only even-pair partners
    RotAngles<M> eangA, eangB;
    int pcA = -1, pcB = -1;
    bool haveA = false, haveB = false;
    double kcurA = 0.0, kcurB = 0.0;  // OBS: this step's own half of the even rotations (applied after the observation)
    if (LENSTEP) {
        double vec[M];
        load_vec<M>(vec, p.vec, T, t, ok);
        int ci;
        haveA = even_partner(p, l0, pcA, ci);
        if (haveA) eangA = rot_angles<M>(vec, (OBS ? sb : sa + sb) * p.cl[ci]), kcurA = sa * p.cl[ci];
        if (pair) {
            haveB = even_partner(p, l0 + 1, pcB, ci);
            if (haveB) eangB = rot_angles<M>(vec, (OBS ? sb : sa + sb) * p.cl[ci]), kcurB = sa * p.cl[ci];
        }
    }
    if (PROG == PROG_H2_CN_H2 && pair) {
        double zv[M];
        load_vec<M>(zv, p.zvec, T, t, ok);
        pang = rpair_angles<M>(zv, ok ? p.zprev[t] : 0.0, sa * p.cl2[p.l_begin + l0]);
    }
    pdl_wait();
    load_rows<M>(A, base, T, t, ok);
    if (pair) load_rows<M>(B, base + chan, T, t, ok);
    if constexpr (OBS) {  // see the pair path above
        double mk[M], vec[M];
#pragma unroll
        for (int k = 0; k < M; ++k) mk[k] = 1.0;
        if (p.flags & F_MASK) load_vec<M>(mk, p.mask, T, t, ok);
        load_vec<M>(vec, p.vec, T, t, ok);
        double *osm = reinterpret_cast<double *>(xs);
        cplx Q[M];
#pragma unroll
        for (int k = 0; k < M; ++k) Q[k] = c_zero();
        if (haveA) {
            load_rows<M>(Q, base + ((ptrdiff_t)pcA - l0) * (ptrdiff_t)chan, T, t, ok);
            rotate_both<M>(A, Q, eangA);
        }
#pragma unroll
        for (int k = 0; k < M; ++k) A[k] = c_scale(A[k], mk[k]), Q[k] = c_scale(Q[k], mk[k]);
        obs_channel<M>(p, A, l0, b, t, mine, tl, Tc, osm);
        if (haveA) rotate_member<M>(A, Q, rot_angles<M>(vec, kcurA));
        if (pair) {
#pragma unroll
            for (int k = 0; k < M; ++k) Q[k] = c_zero();
            if (haveB) {
                load_rows<M>(Q, base + ((ptrdiff_t)pcB - l0) * (ptrdiff_t)chan, T, t, ok);
                rotate_both<M>(B, Q, eangB);
            }
#pragma unroll
            for (int k = 0; k < M; ++k) B[k] = c_scale(B[k], mk[k]), Q[k] = c_scale(Q[k], mk[k]);
            obs_channel<M>(p, B, l0 + 1, b, t, mine, tl, Tc, osm);
            if (haveB) rotate_member<M>(B, Q, rot_angles<M>(vec, kcurB));
        }
    } else if (LENSTEP) {
        cplx Q[M];
        if (haveA) {
            load_rows<M>(Q, base + ((ptrdiff_t)pcA - l0) * (ptrdiff_t)chan, T, t, ok);
            rotate_member<M>(A, Q, eangA);
        }
        if (haveB) {
            load_rows<M>(Q, base + ((ptrdiff_t)pcB - l0) * (ptrdiff_t)chan, T, t, ok);
            rotate_member<M>(B, Q, eangB);
        }
        if (p.flags & F_MASK) {
            double mk[M];
            load_vec<M>(mk, p.mask, T, t, ok);
#pragma unroll
            for (int k = 0; k < M; ++k) {
                A[k] = c_scale(A[k], mk[k]);
                if (pair) B[k] = c_scale(B[k], mk[k]);
            }
        }
    }

    if (PROG == PROG_ROT_CN_ROT || LENSTEP) {
        const RotAngles<M> &ang = rang;  // reused after the CN
        if (pair) rotate_pair<M, false>(A, B, ang);
        cn_channel<M>(A, fA, toff, toff_prev, tl, Tc, sm_scan, short_scan);
        if (pair) {
            cn_channel<M>(B, fB, toff, toff_prev, tl, Tc, sm_scan + 128, short_scan);
            rotate_pair<M, false>(A, B, ang);
        }
    } else if (PROG == PROG_H2_CN_H2) {
        const RPairAngles<M> &ang = pang;  // reused after the CN
        if (pair) h2_pair<M>(A, B, ang, false, tl, Tc, xs);  // (oe, oo)
        cn_channel<M>(A, fA, toff, toff_prev, tl, Tc, sm_scan, short_scan);
        if (pair) {
            cn_channel<M>(B, fB, toff, toff_prev, tl, Tc, sm_scan + 128, short_scan);
            h2_pair<M>(A, B, ang, true, tl, Tc, xs);  // (oo, oe)
        }
    } else if (PROG == PROG_CN) {
        cplx A0[M];
#pragma unroll
        for (int k = 0; k < M; ++k) A0[k] = A[k];
        cn_channel<M>(A, fA, toff, toff_prev, tl, Tc, sm_scan, short_scan);
        if (p.flags & F_SOLVE_ONLY) {  // cn_channel returns 2x - g with x = (1 + i tau H0)^-1 g
#pragma unroll
            for (int k = 0; k < M; ++k) A[k] = c_make(0.5 * (A[k].x + A0[k].x), 0.5 * (A[k].y + A0[k].y));
        }
        if (p.flags & F_MASK) {
            double mk[M];
            load_vec<M>(mk, p.mask, T, t, ok);
#pragma unroll
            for (int k = 0; k < M; ++k) A[k] = c_scale(A[k], mk[k]);
        }
    } else if (PROG == PROG_LINE_SO_LEN) {
        // P = exp(-i tau (-q z E)) = exp(-i s w_z)   mesh_operators.py:329-341
        double vec[M];
        load_vec<M>(vec, p.vec, T, t, ok);
        cplx ph[M];
        {
            double th[M], sn[M], cs[M];
#pragma unroll
            for (int k = 0; k < M; ++k) th[k] = sa * vec[k];
            fast_sincos_n<M>(th, sn, cs);
#pragma unroll
            for (int k = 0; k < M; ++k) {
                ph[k] = c_make(cs[k], -sn[k]);
                A[k] = c_mul(ph[k], A[k]);
            }
        }
        cn_channel<M>(A, fA, toff, toff_prev, tl, Tc, sm_scan, short_scan);
#pragma unroll
        for (int k = 0; k < M; ++k) A[k] = c_mul(ph[k], A[k]);
        if (p.flags & F_MASK) {
            double mk[M];
            load_vec<M>(mk, p.mask, T, t, ok);
#pragma unroll
            for (int k = 0; k < M; ++k) A[k] = c_scale(A[k], mk[k]);
        }
    } else if (PROG == PROG_LINE_CN) {
        double wz[M];
        load_vec<M>(wz, p.vec, T, t, ok);
        cplx D[M];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            cplx th = ok ? ld_c(p.th + k * T + t) : c_zero();
            D[k] = c_make(1.0 - th.y, fma(sa, wz[k], th.x));  // 1 + i (tau h + tau E w_z)
        }
        line_cn_factors<M>(fA, D, toff, toff_prev, tl, xs);
        cn_channel<M>(A, fA, toff, toff_prev, tl, Tc, sm_scan, short_scan > 0 ? short_scan : 1);
        if (p.flags & F_MASK) {
            double mk[M];
            load_vec<M>(mk, p.mask, T, t, ok);
#pragma unroll
            for (int k = 0; k < M; ++k) A[k] = c_scale(A[k], mk[k]);
        }
    } else if (PROG == PROG_LINE_SO_VEL) {
        // theta identical for every z-pair: zvec holds v_pref on rows that start a pair  mesh_operators.py:384-427
        double zv[M];
        load_vec<M>(zv, p.zvec, T, t, ok);
        const RPairAngles<M> ang = rpair_angles<M>(zv, ok ? p.zprev[t] : 0.0, sa);
        rpair_layer_even<M, false>(A, A, ang);
        rpair_layer_odd<M, false>(A, A, ang, tl, Tc, xs);
        cn_channel<M>(A, fA, toff, toff_prev, tl, Tc, sm_scan, short_scan);
        rpair_layer_odd<M, false>(A, A, ang, tl, Tc, xs);
        rpair_layer_even<M, false>(A, A, ang);
        if (p.flags & F_MASK) {
            double mk[M];
            load_vec<M>(mk, p.mask, T, t, ok);
#pragma unroll
            for (int k = 0; k < M; ++k) A[k] = c_scale(A[k], mk[k]);
        }
    }
    store_rows<M>(A, obase, T, t, mine);
    if (pair) store_rows<M>(B, obase + chan, T, t, mine);
}

// ---------------------------------------------------------------------------------------------
// Generic l-sweep for an ODD number of channels, where the reference's even/odd membership follows the
// parity of the flat index j*L + l and therefore alternates with j (mesh_operators.py:1045).
// One thread per (pair, position); pairs of one sweep are disjoint at fixed j.
// ---------------------------------------------------------------------------------------------
struct SweepParams {
    cplx *psi;
    const double *vec;
    const double *cl;
    const double *mask;
    const double *scal_a;
    const double *scal_b;
    int L, L_total, l_begin, T, M, R;
    int parity;
    int flags;
};

__global__ void k_sweep_flat(const SweepParams p)
{
    const int Rp = p.M * p.T;
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= Rp) return;
    const int l = blockIdx.y;  // local pair index: channels (l, l+1)
    const int b = blockIdx.z;
    const int k = pos / p.T, t = pos % p.T;
    const long long i = (long long)t * p.M + k;
    if (i >= p.R) return;
    const int gl = p.l_begin + l;
    if ((int)((i * (long long)p.L_total + gl) & 1) != p.parity) return;
    const double s = (p.scal_a ? p.scal_a[b] : 0.0) + (p.scal_b ? p.scal_b[b] : 0.0);
    cplx *pa = p.psi + ((size_t)b * p.L + l) * Rp + pos;
    cplx *pb = pa + Rp;
    cplx a = *pa, bb = *pb;
    double sn, cs;
    sincos(s * p.cl[gl] * p.vec[pos], &sn, &cs);
    if (p.flags & F_REAL_ROT) {
        *pa = c_make(fma(cs, a.x, sn * bb.x), fma(cs, a.y, sn * bb.y));
        *pb = c_make(fma(cs, bb.x, -sn * a.x), fma(cs, bb.y, -sn * a.y));
    } else {
        *pa = c_make(fma(cs, a.x, sn * bb.y), fma(cs, a.y, -sn * bb.x));
        *pb = c_make(fma(cs, bb.x, sn * a.y), fma(cs, bb.y, -sn * a.x));
    }
}

// point-wise mask over everything (generic path)
__global__ void k_mask(cplx *psi, const double *mask, int Rp, long long n_channels)
{
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= Rp) return;
    const double m = mask[pos];
    for (long long c = blockIdx.y; c < n_channels; c += gridDim.y) {
        cplx *q = psi + (size_t)c * Rp + pos;
        *q = c_scale(*q, m);
    }
}

// ---------------------------------------------------------------------------------------------
// LU factors of (1 + i tau H0) per channel, written in the interleaved layout.  One thread per channel
// (serial Thomas elimination, run once per distinct tau -- the matrices are time-independent for
// SphericalHarmonicMesh).  h_diag: [L][R] reference layout.
// ---------------------------------------------------------------------------------------------
__global__ void k_factor(const cplx *__restrict__ h_diag, const double *__restrict__ h_off, double tau, int L, int R,
                         int M, int T, cplx *__restrict__ w)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int Rp = M * T;
    cplx *wl = w + (size_t)l * Rp;
    cplx wprev = c_zero();
    for (int i = 0; i < Rp; ++i) {
        cplx wi = c_make(1.0, 0.0);
        if (i < R) {
            cplx h = h_diag[(size_t)l * R + i];
            cplx piv = c_make(1.0 - tau * h.y, tau * h.x);  // 1 + i tau h
            if (i > 0) {
                double o = tau * h_off[i - 1];               // O = i*o,  O^2 = -o^2
                piv = c_make(fma(o * o, wprev.x, piv.x), fma(o * o, wprev.y, piv.y));
            }
            wi = c_inv(piv);
            wprev = wi;
        }
        wl[(size_t)(i % M) * T + (i / M)] = wi;
    }
}

// chunk aggregates P_t = prod_{i=tM-1}^{tM+M-2} e_i, Q_t = prod_{i=tM}^{tM+M-1} e_i, e_i = -i tau off_i w_i
__global__ void k_aggregates(const cplx *__restrict__ w, const double *__restrict__ toff, int L, int M, int T,
                             cplx *__restrict__ aggP, cplx *__restrict__ aggQ)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (t >= T) return;
    const cplx *wl = w + (size_t)l * M * T;
    cplx P, Q = c_make(1.0, 0.0);
    if (t == 0) {
        P = c_zero();
    } else {
        cplx wp = wl[(size_t)(M - 1) * T + t - 1];
        double tp = toff[(size_t)(M - 1) * T + t - 1];
        P = c_make(tp * wp.y, -tp * wp.x);
    }
    for (int k = 0; k < M; ++k) {
        cplx wk = wl[(size_t)k * T + t];
        double tk = toff[(size_t)k * T + t];
        cplx e = c_make(tk * wk.y, -tk * wk.x);
        if (k < M - 1) P = c_mul(P, e);
        Q = c_mul(Q, e);
    }
    aggP[(size_t)l * T + t] = P;
    aggQ[(size_t)l * T + t] = Q;
}

// tau * h_diag of a single channel, permuted (PROG_LINE_CN)
__global__ void k_make_th(const cplx *__restrict__ h_diag, double tau, int R, int M, int T, cplx *__restrict__ th)
{
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= M * T) return;
    const int k = pos / T, t = pos % T;
    const long long i = (long long)t * M + k;
    th[pos] = (i < R) ? c_scale(h_diag[i], tau) : c_zero();
}

// log-magnitude of the product of the chunk multipliers over each group of G consecutive threads (G = 32: a warp): the host
// takes the maximum to decide whether the cross-warp part of the CN scans is short-ranged (common.cuh), and whether a
// half-warp halo (G = 16) is enough for the r-segments of the length gauge.
__global__ void k_scan_bound(const cplx *__restrict__ aggP, const cplx *__restrict__ aggQ, int L, int T, int G, double *__restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (l, group)
    const int nw = T / G;
    if (idx >= L * nw) return;
    const int l = idx / nw, w = idx % nw;
    double sp = 0.0, sq = 0.0;
    for (int t = w * G; t < w * G + G; ++t) {
        sp += 0.5 * log(fmax(c_abs2(aggP[(size_t)l * T + t]), 1e-300));
        sq += 0.5 * log(fmax(c_abs2(aggQ[(size_t)l * T + t]), 1e-300));
    }
    out[idx] = fmax(sp, sq);
}

// ---------------------------------------------------------------------------------------------
// layout conversion  reference [n][R]  <->  interleaved [n][M][T]
// ---------------------------------------------------------------------------------------------
// src_period < n: the source holds src_period channels that are repeated (one member's g broadcast to every member of an ensemble)
__global__ void k_to_internal_c(const cplx *__restrict__ src, cplx *__restrict__ dst, int R, int M, int T, long long n, long long src_period)
{
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const int Rp = M * T;
    if (pos >= Rp) return;
    const int k = pos / T, t = pos % T;
    const long long i = (long long)t * M + k;
    for (long long c = blockIdx.y; c < n; c += gridDim.y)
        dst[(size_t)c * Rp + pos] = (i < R) ? src[(size_t)(c % src_period) * R + i] : c_zero();
}
__global__ void k_from_internal_c(const cplx *__restrict__ src, cplx *__restrict__ dst, int R, int M, int T, long long n)
{
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const int Rp = M * T;
    if (pos >= Rp) return;
    const int k = pos / T, t = pos % T;
    const long long i = (long long)t * M + k;
    if (i >= R) return;
    for (long long c = blockIdx.y; c < n; c += gridDim.y) dst[(size_t)c * R + i] = src[(size_t)c * Rp + pos];
}

// ---------------------------------------------------------------------------------------------
// Observables (kernel 4).  Stage 1: one CTA per (channel, sim) reduces its row(s) to a few partial sums;
// stage 2 adds the per-channel partials in a fixed order (deterministic, no atomics).
// partial layout per (b, l): [norm_l, r_l, z_l, h0_l, within_0..within_{nr-1}]
// ---------------------------------------------------------------------------------------------
#ifndef ION_MAX_RADII
#define ION_MAX_RADII 8
#endif
struct ObserveParams {
    const cplx *psi;          // [batch][L][Rp]
    const double *rvec;       // [Rp] permuted r_j (0 padding)
    const cplx *h_diag;       // [L][R] reference layout (or nullptr)
    const double *h_off;      // [R-1]
    const double *cl_z;       // [L_total-1] c_l (for <z>), or nullptr
    const cplx *state_rows;   // [n_states][Rp] permuted
    const int *state_first;   // [L+1] CSR offsets: states of local channel l are state_order[first[l]..first[l+1])
    const int *state_order;   // [n_states]
    double *partial;          // [batch][L][4 + nr]
    double *ip_out;           // [batch][n_states][2]
    double radii[ION_MAX_RADII];
    int n_radii;
    int n_states;
    int L, L_total, l_begin, R, M, T;
    unsigned what;
    double ipm;
    int line;                 // LineMesh: <z> = sum z |g|^2 is the "r" observable; no l coupling
    int ghost_hi;             // an upper ghost channel follows the last owned channel (l-block shard): <z> couples to it
};

__global__ void __launch_bounds__(256) k_observe(const ObserveParams p)
{
    __shared__ double sm[32 * (4 + ION_MAX_RADII)];
    const int l = blockIdx.x, b = blockIdx.y;
    const int Rp = p.M * p.T;
    const cplx *row = p.psi + ((size_t)b * p.L + l) * Rp;
    const bool has_up = (l + 1 < p.L) || p.ghost_hi;
    const bool want_z = (p.what & 16u) && has_up && !p.line && p.cl_z;
    const bool want_h = (p.what & 32u) && p.h_diag;
    double acc[4 + ION_MAX_RADII];
#pragma unroll
    for (int q = 0; q < 4 + ION_MAX_RADII; ++q) acc[q] = 0.0;
    for (int pos = threadIdx.x; pos < Rp; pos += blockDim.x) {
        const int k = pos / p.T, t = pos % p.T;
        const int i = t * p.M + k;
        if (i >= p.R) continue;
        cplx g = row[pos];
        double n2 = c_abs2(g);
        acc[0] += n2;
        double r = p.rvec ? p.rvec[pos] : 0.0;
        acc[1] += r * n2;
        if (want_z) {
            cplx gu = row[Rp + pos];
            acc[2] += 2.0 * (g.x * gu.x + g.y * gu.y) * r;
        }
        if (want_h) {
            cplx h = p.h_diag[(size_t)l * p.R + i];
            double v = h.x * n2;
            if (i + 1 < p.R) {
                int i1 = i + 1;
                cplx gn = row[(size_t)(i1 % p.M) * p.T + (i1 / p.M)];
                v += 2.0 * p.h_off[i] * (g.x * gn.x + g.y * gn.y);
            }
            acc[3] += v;
        }
        for (int q = 0; q < p.n_radii; ++q)
            if (r <= p.radii[q]) acc[4 + q] += n2;
    }
    if (want_z) acc[2] *= p.cl_z[p.l_begin + l];
    block_sum<4 + ION_MAX_RADII>(acc, sm, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        double *out = p.partial + ((size_t)b * p.L + l) * (4 + p.n_radii);
        out[0] = acc[0];
        out[1] = acc[1];
        out[2] = acc[2];
        out[3] = acc[3];
        for (int q = 0; q < p.n_radii; ++q) out[4 + q] = acc[4 + q];
    }
    // inner products with the test states living in this channel (mesh/meshes.py:1117-1129)
    if ((p.what & 2u) && p.n_states > 0) {
        for (int si = p.state_first[l]; si < p.state_first[l + 1]; ++si) {
            const int s = p.state_order[si];
            const cplx *sr = p.state_rows + (size_t)s * Rp;
            double ip[2] = {0.0, 0.0};
            for (int pos = threadIdx.x; pos < Rp; pos += blockDim.x) {
                cplx a = sr[pos], g = row[pos];  // conj(a) * g ; padding rows are zero in both
                ip[0] += a.x * g.x + a.y * g.y;
                ip[1] += a.x * g.y - a.y * g.x;
            }
            block_sum<2>(ip, sm, threadIdx.x, blockDim.x);
            if (threadIdx.x == 0) {
                p.ip_out[((size_t)b * p.n_states + s) * 2 + 0] = ip[0] * p.ipm;
                p.ip_out[((size_t)b * p.n_states + s) * 2 + 1] = ip[1] * p.ipm;
            }
        }
    }
}

// stage 2: record assembly.  One CTA per simulation; the sums over the channels are strided per thread and then combined by
// block_sum's fixed tree, so the result does not depend on scheduling (deterministic, no atomics).
__global__ void __launch_bounds__(256) k_observe_finish(const double *__restrict__ partial, const double *__restrict__ ip, double *__restrict__ out,
                                                         int batch, int L, int n_states, int n_radii, unsigned what, double ipm, long long rec)
{
    __shared__ double sm[32 * (4 + ION_MAX_RADII)];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int np = 4 + n_radii;
    pdl_launch_dependents();  // nobody downstream reads the record: the next step's kernels need not wait for it
    pdl_wait();
    const double *pb = partial + (size_t)b * L * np;
    double *o = out + (size_t)b * rec;
    double sums[4 + ION_MAX_RADII];
#pragma unroll
    for (int q = 0; q < 4 + ION_MAX_RADII; ++q) sums[q] = 0.0;
    for (int l = tid; l < L; l += blockDim.x) {
#pragma unroll
        for (int q = 0; q < 4 + ION_MAX_RADII; ++q)
            if (q < np) sums[q] += pb[(size_t)l * np + q];
    }
    block_sum<4 + ION_MAX_RADII>(sums, sm, tid, blockDim.x);
    long long c = 0;
    if (what & 1u) {
        if (tid == 0) o[c] = sums[0] * ipm;
        c += 1;
    }
    if (what & 2u) {
        for (int s = tid; s < 2 * n_states; s += blockDim.x) o[c + s] = ip[(size_t)b * n_states * 2 + s];
        c += 2 * n_states;
    }
    if (what & 4u) {
        for (int l = tid; l < L; l += blockDim.x) o[c + l] = fabs(pb[(size_t)l * np] * ipm);
        c += L;
    }
    if (tid == 0) {
        if (what & 8u) o[c++] = sums[1] * ipm;
        if (what & 16u) o[c++] = sums[2] * ipm;
        if (what & 32u) o[c++] = sums[3] * ipm;
        if (what & 64u)
            for (int q = 0; q < n_radii; ++q) o[c++] = sums[4 + q] * ipm;
    }
}

// ---------------------------------------------------------------------------------------------
// ion_tdma_c128: general (non-symmetric, matrix given per call) Thomas solve, cy.pyx:9-50.
// One warp-sized CTA per system: the data is staged through shared memory with coalesced accesses and the
// two recurrences are run by lane 0 exactly as the reference runs them (two divisions per row).
// This entry point exists for drop-in parity with cy.tdma; the production path uses cn_channel.
// ---------------------------------------------------------------------------------------------
__global__ void k_tdma_general(const cplx *__restrict__ sub, const cplx *__restrict__ diag, const cplx *__restrict__ sup,
                               const cplx *__restrict__ rhs, cplx *__restrict__ x, long long n, cplx *scratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long sys = blockIdx.x;
    // c' and d' live in shared memory when they fit, else in a global scratch of 2n per system
    cplx *cp = scratch ? scratch + sys * 2 * n : reinterpret_cast<cplx *>(smem_raw);  // [n]
    cplx *dp = cp + n;                                                                 // [n]
    const cplx *a = sub + sys * (n - 1), *bdi = diag + sys * n, *c = sup + sys * (n - 1), *d = rhs + sys * n;
    // stage sup and rhs (coalesced); the recurrence overwrites them with c' and d'
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        cp[i] = (i < n - 1) ? c[i] : c_zero();
        dp[i] = d[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx inv = c_inv(bdi[0]);
        cp[0] = c_mul(cp[0], inv);
        dp[0] = c_mul(dp[0], inv);
        for (long long i = 1; i < n; ++i) {
            cplx s = a[i - 1];
            cplx denom = c_sub(bdi[i], c_mul(s, cp[i - 1]));
            inv = c_inv(denom);
            cp[i] = c_mul(cp[i], inv);
            dp[i] = c_mul(c_sub(dp[i], c_mul(s, dp[i - 1])), inv);
        }
        for (long long i = n - 2; i >= 0; --i) dp[i] = c_sub(dp[i], c_mul(cp[i], dp[i + 1]));
    }
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += blockDim.x) x[sys * n + i] = dp[i];
}

}  // namespace ion
