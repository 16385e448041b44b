// ionization_b200 -- the folded length-gauge step with FOUR points per thread (sm_100a).
//
// k_unit<PROG_LEN_STEP> (kernels.cuh) gives a thread 4 rows of BOTH channels of its odd pair (8 complex values, 128 registers,
// 16 warps per SM): every phase of a CTA -- coefficient loads, psi loads, the two scans of the solve, stores -- is exposed,
// and a time step of one simulation is two consecutive CTA lifetimes (251 units on 148 SMs).  Here a thread owns the 4 rows
// 4p .. 4p+3 of ONE channel (even lanes: the lower channel of the pair, odd lanes: the upper one), i.e. 4 complex values in
// <= 64 registers: a pair is a CTA of 2T <= 1024 threads, 32 warps per SM (8 per scheduler) hide the latencies that 4 could
// not, and a thread's serial work is halved.  Same arithmetic per operator as PROG_LEN_STEP:
//     even rotation by (s_a + s_b) c x_j with the READ-ONLY even-pair partner [+ mask]      rotate_member
//     odd rotation by s_a c x_j with the other channel of the pair (lane ^ 1: 4 values by shuffle)
//     Crank-Nicolson on the own channel: 2 (1 + i tau H0)^-1 g - g
//     odd rotation again; out of place.
// Crank-Nicolson with the recurrences written for u = w y (one real-by-complex and one complex multiplication per row
// instead of two complex ones):  forward  y_k = g_k - i toff_{k-1} u_{k-1},  u_k = w_k y_k;
//                                backward x_k = u_k + w_k (-i toff_k x_{k+1});  out_k = 2 x_k - g_k.
// Zero-inflow pass, scan of the chunk maps over the 16 lanes of the same channel (the multipliers are k_aggregates' P, Q: the
// forward inflow is y of the previous row, the backward inflow x of the next one), true-inflow pass -- as cn_channel.
#pragma once
#include "kernels.cuh"

namespace ion {

// -i * t * v
ION_DEVINL cplx mi_scale(double t, cplx v) { return c_make(t * v.y, -t * v.x); }

// cos/sin of theta_k = sc * vec[k], k = 0..3, as a "core": (c, s) of the first row and of the increment per row when vec is linear in
// the row index (vec_dv != 0: the rows follow by angle addition while they are being used, so only four numbers stay live);
// otherwise the core holds sc itself and the rows are evaluated one by one
struct RotCore {
    double c0, s0, cd, sd;
};
ION_DEVINL RotCore rot_core(double v0, double vec_dv, double sc)
{
    RotCore r;
    if (vec_dv != 0.0) {
        const double th[2] = {sc * v0, sc * vec_dv};
        double sn[2], cs[2];
        fast_sincos_n<2>(th, sn, cs);
        r.c0 = cs[0], r.s0 = sn[0], r.cd = cs[1], r.sd = sn[1];
    } else {
        r.c0 = sc, r.s0 = 0.0, r.cd = 0.0, r.sd = 0.0;
    }
    return r;
}
// X <- cos(theta) X - i sin(theta) Y, row by row  (one member of the symmetric rotation [[c, -i s], [-i s, c]])
ION_DEVINL void rotate_member_core(cplx (&X)[4], const cplx (&Y)[4], const RotCore &r, double vec_dv, const double *__restrict__ vecp, int T, int pp)
{
    if (vec_dv != 0.0) {
        double c = r.c0, sn = r.s0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const cplx x = X[k], q = Y[k];
            X[k] = c_make(fma(c, x.x, sn * q.y), fma(c, x.y, -sn * q.x));
            if (k < 3) {
                const double c1 = fma(c, r.cd, -sn * r.sd);
                sn = fma(sn, r.cd, c * r.sd);
                c = c1;
            }
        }
    } else {
        double th[4], sn[4], cs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) th[k] = r.c0 * vecp[k * T + pp];
        fast_sincos_n<4>(th, sn, cs);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const cplx x = X[k], q = Y[k];
            X[k] = c_make(fma(cs[k], x.x, sn[k] * q.y), fma(cs[k], x.y, -sn[k] * q.x));
        }
    }
}

// dynamic shared memory of a CTA of Tc threads: scan scratch 256 cplx | g stash 4 Tc cplx | rotation core 2 Tc cplx
inline size_t len4_smem_bytes(int Tc) { return (256 + 6 * (size_t)Tc) * sizeof(cplx); }

template <int TMAX>
__global__ void __launch_bounds__(TMAX, 1024 / TMAX) k_len4(const UnitParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm = reinterpret_cast<cplx *>(smem_raw);  // 4 x 64 cplx: (P, B) of the warp aggregates, forward and backward scans
    const int tl = threadIdx.x, Tc = blockDim.x, T = p.T;
    cplx *gst = sm + 256 + tl;                      // my column of the g stash: row k at gst[k * Tc]
    cplx *cst = sm + 256 + 4 * (size_t)Tc + 2 * tl; // my odd-rotation core

    const int ch = tl & 1, pp = tl >> 1;
    const int unit = p.unit0 + (int)blockIdx.x * p.unit_stride, b = blockIdx.y;
    int l0;
    bool pair = true;
    unit_channels(p, unit, l0, pair);
    const int lc = l0 + (pair ? ch : 0);  // my channel; a single channel is processed by both lanes of a pair alike (odd lanes do not store)
    const size_t chan = (size_t)4 * T;
    const double sa = p.scal_a ? p.scal_a[b] : 0.0, sb = p.scal_b ? p.scal_b[b] : 0.0;
    int pc, ci;
    const bool have_partner = even_partner(p, lc, pc, ci);

    pdl_launch_dependents();
    // ---- psi-independent prologue: the two rotation cores ----
    const double v0 = p.vec[pp];
    const RotCore ce = rot_core(v0, p.vec_dv, have_partner ? (sa + sb) * p.cl[ci] : 0.0);
    {
        const RotCore co = rot_core(v0, p.vec_dv, pair ? sa * p.cl[p.l_begin + l0] : 0.0);
        cst[0] = c_make(co.c0, co.s0);
        cst[1] = c_make(co.cd, co.sd);
    }

    pdl_wait();
    cplx X[4];
    load_rows<4>(X, p.psi + ((size_t)b * p.L + lc) * chan, T, pp, true);
    if (have_partner) {
        cplx Qp[4];
        load_rows<4>(Qp, p.psi + ((size_t)b * p.L + pc) * chan, T, pp, true);
        rotate_member_core(X, Qp, ce, p.vec_dv, p.vec, T, pp);
    }
    if (p.flags & F_MASK) {
#pragma unroll
        for (int k = 0; k < 4; ++k) X[k] = c_scale(X[k], p.mask[k * T + pp]);
    }

    // ---- odd rotation with the other channel of the pair ----
    if (pair) {
        cplx Y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) Y[k] = make_double2(__shfl_xor_sync(0xffffffffu, X[k].x, 1), __shfl_xor_sync(0xffffffffu, X[k].y, 1));
        RotCore co;
        co.c0 = cst[0].x, co.s0 = cst[0].y, co.cd = cst[1].x, co.sd = cst[1].y;
        rotate_member_core(X, Y, co, p.vec_dv, p.vec, T, pp);
    }

    // ---- Crank-Nicolson on my channel ----
    {
        const cplx *wch = p.w + (size_t)lc * chan;
        cplx w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = ld_c(wch + k * T + pp);
        double to[4];
        load_vec<4>(to, p.toff, T, pp, true);
        const int reach = p.short_scan > 0 ? 2 * p.short_scan : 0;  // a warp spans 64 rows of a channel here, 128 in k_scan_bound's layout
        // forward, zero inflow: only y of the last row leaves the chunk.  g is parked in shared memory for the later passes.
        cplx y = X[0], u = c_mul(w[0], y);
        gst[0] = X[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            gst[k * Tc] = X[k];
            y = c_add(X[k], mi_scale(to[k - 1], u));
            if (k < 3) u = c_mul(w[k], y);
        }
        const cplx yin = affine_scan_strided_exclusive<true, 2>(ld_c(p.aggP + (size_t)lc * T + pp), y, sm, sm + 64, tl, Tc, reach);
        // forward, true inflow
        cplx U[4];
        {
            const cplx wprev = pp > 0 ? ld_c(wch + 3 * T + pp - 1) : c_zero();
            y = c_fma(mi_scale(p.toff_prev[pp], wprev), yin, gst[0]);
        }
        U[0] = c_mul(w[0], y);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            y = c_add(gst[k * Tc], mi_scale(to[k - 1], U[k - 1]));
            U[k] = c_mul(w[k], y);
        }
        // backward, zero inflow
        cplx x = U[3];
#pragma unroll
        for (int k = 2; k >= 0; --k) x = c_fma(w[k], mi_scale(to[k], x), U[k]);
        const cplx xin = affine_scan_strided_exclusive<false, 2>(ld_c(p.aggQ + (size_t)lc * T + pp), x, sm + 128, sm + 192, tl, Tc, reach);
        // backward, true inflow; out = 2 x - g
        x = xin;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            x = c_fma(w[k], mi_scale(to[k], x), U[k]);
            const cplx g = gst[k * Tc];
            X[k] = c_make(fma(2.0, x.x, -g.x), fma(2.0, x.y, -g.y));
        }
    }

    if (pair) {
        cplx Y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) Y[k] = make_double2(__shfl_xor_sync(0xffffffffu, X[k].x, 1), __shfl_xor_sync(0xffffffffu, X[k].y, 1));
        RotCore co;
        co.c0 = cst[0].x, co.s0 = cst[0].y, co.cd = cst[1].x, co.sd = cst[1].y;
        rotate_member_core(X, Y, co, p.vec_dv, p.vec, T, pp);
    }
    if (pair || ch == 0) store_rows<4>(X, p.psi_out + ((size_t)b * p.L + lc) * chan, T, pp, true);
}

}  // namespace ion
