// ionization_b200 -- SphericalHarmonicMesh length gauge, AlternatingDirectionImplicit (ION_SH_LEN_ADI).
//
// Reference: evolution_methods.py:49-77 with total_hamiltonian = [H0 (r-wrapped), H_int (l-wrapped)]
// (mesh_operators.py:1020-1035), i.e. per time step
//
//      g <- mask * (1 + i tau H0)^-1_r  (1 - i tau H_int)_l  (1 + i tau H_int)^-1_l  (1 - i tau H0)_r  g
//
// H_int couples channels l <-> l+1 at radius j with  E(t) * c_l * x_j  (x_j = -q r_j), zero diagonal
// (mesh_operators.py:988-1018): for every radial point an L x L tridiagonal system ALONG l.
//
// k_adi_l does the first three operators in one out-of-place pass over psi:
//   * (1 - i tau H0)_r is applied while loading: three reads per point (own row and both r-neighbours, which sit at
//     other positions of the row-interleaved layout and mostly hit L1/L2);
//   * the l-solve: a CTA owns PW consecutive positions x all L channels; thread (c, q) holds the CL = 8 channels
//     8c .. 8c+7 of position q in registers.  With s = tau E x_j and beta_l = s c_l the matrix is 1 on the diagonal and
//     i beta_l off it, so its pivots are REAL: p_l = 1 + beta_{l-1}^2 / p_{l-1}.  w_l = 1/p_l obeys the Moebius map
//     w -> 1 / (1 + a w); every thread composes its 8 maps into one 2x2 real matrix, the chunk inflows follow from a
//     two-level prefix (shuffles over the chunks a warp holds, then a walk over the warp totals in shared memory), exactly.
//     Forward / backward substitution are affine recurrences with multipliers -i beta w: zero-inflow pass, two-level prefix
//     of the chunk aggregates, second pass with the true inflow -- the same scheme as the radial Crank-Nicolson scans
//     (kernels.cuh), across chunks instead of lanes;
//   * (1 - i tau H_int)(1 + i tau H_int)^-1 v = 2 (1 + i tau H_int)^-1 v - v: no second mat-vec.
// The remaining (1 + i tau H0)^-1_r and the mask are k_unit<PROG_CN> with F_SOLVE_ONLY | F_MASK.
// PW = 8 positions are one full 128-byte line per channel; L <= 8 * 512 channels.
#pragma once
#include "common.cuh"
#include "kernels.cuh"  // e_of

namespace ion {

struct AdiParams {
    const cplx *psi;         // [batch][L][M][T]
    cplx *out;               // same shape, a different buffer
    const cplx *thd;         // [L][M][T]  tau * h_diag, permuted (k_make_thd)
    const double *toff;      // [M][T]     tau * h_off[i] (0 for i >= R-1)
    const double *toff_prev; // [T]        tau * h_off[t*M - 1] (0 for t = 0)
    const double *vec;       // [M][T]     x_j, 0 in the padding
    const double *cl;        // [L-1]      c_l
    const double *scal;      // [batch]    tau * E of this step
    int L, T, M, PW, NC;
};

constexpr int ADI_CL = 8;
constexpr int ADI_MAX_THREADS = 512;

// (-i e) v for real e
ION_DEVINL cplx mul_mi(double e, cplx v) { return c_make(e * v.y, -e * v.x); }
// (-i e) v + a
ION_DEVINL cplx fma_mi(double e, cplx v, cplx a) { return c_make(fma(e, v.y, a.x), fma(-e, v.x, a.y)); }

// ---- two-level prefixes over the chunks of one radial position ----------------------------------------------------
// Threads of the same position q sit PW lanes apart; a warp holds CPW = 32 / PW consecutive chunks of each of its positions.
// Level 1: Kogge-Stone over those chunks with shuffles (stride PW).  Level 2: the warp totals go to shared memory and every
// thread walks over the totals of the preceding (following) warps -- at most 15 steps instead of one per chunk.

// affine maps v -> M v + Y.  Returns the value entering this thread's chunk; sm: 4 * nw * PW doubles.
template <bool FWD>
ION_DEVINL cplx adi_affine_inflow(cplx M, cplx Y, double *sm, int PW, int tid, int nthreads)
{
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5, q = lane % PW, stride = nw * PW;
    for (int s = PW; s < 32; s <<= 1) {
        const cplx Mp = FWD ? shfl_up_c(M, s) : shfl_down_c(M, s);
        const cplx Yp = FWD ? shfl_up_c(Y, s) : shfl_down_c(Y, s);
        if (FWD ? (lane >= s) : (lane + s < 32)) {
            Y = c_fma(M, Yp, Y);
            M = c_mul(M, Mp);
        }
    }
    if (FWD ? (lane >= 32 - PW) : (lane < PW)) {
        const int j = warp * PW + q;
        sm[0 * stride + j] = M.x;
        sm[1 * stride + j] = M.y;
        sm[2 * stride + j] = Y.x;
        sm[3 * stride + j] = Y.y;
    }
    __syncthreads();
    cplx v = c_zero();
    if (FWD) {
        for (int w = 0; w < warp; ++w) {
            const int j = w * PW + q;
            v = c_fma(c_make(sm[0 * stride + j], sm[1 * stride + j]), v, c_make(sm[2 * stride + j], sm[3 * stride + j]));
        }
    } else {
        for (int w = nw - 1; w > warp; --w) {
            const int j = w * PW + q;
            v = c_fma(c_make(sm[0 * stride + j], sm[1 * stride + j]), v, c_make(sm[2 * stride + j], sm[3 * stride + j]));
        }
    }
    __syncthreads();  // sm is reused by the next phase
    // exclusive inside the warp: the inclusive map of the neighbouring chunk applied to the warp's inflow
    const cplx Me = FWD ? shfl_up_c(M, PW) : shfl_down_c(M, PW);
    const cplx Ye = FWD ? shfl_up_c(Y, PW) : shfl_down_c(Y, PW);
    const bool first = FWD ? (lane < PW) : (lane >= 32 - PW);
    return first ? v : c_fma(Me, v, Ye);
}

// Moebius maps w -> (A w + B) / (C w + D) with non-negative real entries (forward only).  Products are renormalised (a Moebius map
// is projective), so long runs of channels cannot overflow.  Returns the value entering this thread's chunk (start value 1).
ION_DEVINL double adi_moebius_inflow(double A, double B, double C, double D, double *sm, int PW, int tid, int nthreads)
{
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5, q = lane % PW, stride = nw * PW;
    for (int s = PW; s < 32; s <<= 1) {
        const double Ap = __shfl_up_sync(0xffffffffu, A, s), Bp = __shfl_up_sync(0xffffffffu, B, s);
        const double Cp = __shfl_up_sync(0xffffffffu, C, s), Dp = __shfl_up_sync(0xffffffffu, D, s);
        if (lane >= s) {  // mine * previous
            const double nA = fma(A, Ap, B * Cp), nB = fma(A, Bp, B * Dp), nC = fma(C, Ap, D * Cp), nD = fma(C, Bp, D * Dp);
            const double sc = 1.0 / (nA + nB + nC + nD);
            A = nA * sc, B = nB * sc, C = nC * sc, D = nD * sc;
        }
    }
    if (lane >= 32 - PW) {
        const int j = warp * PW + q;
        sm[0 * stride + j] = A;
        sm[1 * stride + j] = B;
        sm[2 * stride + j] = C;
        sm[3 * stride + j] = D;
    }
    __syncthreads();
    double v = 1.0;
    for (int w = 0; w < warp; ++w) {
        const int j = w * PW + q;
        v = fma(sm[0 * stride + j], v, sm[1 * stride + j]) / fma(sm[2 * stride + j], v, sm[3 * stride + j]);
    }
    __syncthreads();
    const double Ae = __shfl_up_sync(0xffffffffu, A, PW), Be = __shfl_up_sync(0xffffffffu, B, PW);
    const double Ce = __shfl_up_sync(0xffffffffu, C, PW), De = __shfl_up_sync(0xffffffffu, D, PW);
    return lane < PW ? v : fma(Ae, v, Be) / fma(Ce, v, De);
}

// grid = (Rp / PW, batch); block = PW * NC rounded up to whole warps (threads beyond chunk NC - 1 hold no channels)
__global__ void __launch_bounds__(ADI_MAX_THREADS, 1) k_adi_l(const AdiParams p)
{
    constexpr int CL = ADI_CL;
    __shared__ double sm[4 * ADI_MAX_THREADS];
    const int NT = blockDim.x, tid = threadIdx.x;
    const int q = tid % p.PW, c = tid / p.PW;
    const int T = p.T, M = p.M, L = p.L;
    const size_t Rp = (size_t)M * T;
    const int pos = blockIdx.x * p.PW + q;
    const int b = blockIdx.y;
    const int l0 = c * CL;
    pdl_launch_dependents();

    // r-neighbours of row i = t*M + k in the interleaved layout
    const int k_row = pos / T, t = pos % T;
    double t_hi = p.toff[pos], t_lo;
    int pos_lo, pos_hi;
    if (k_row > 0) {
        pos_lo = pos - T;
        t_lo = p.toff[pos_lo];
    } else if (t > 0) {
        pos_lo = (M - 1) * T + t - 1;
        t_lo = p.toff_prev[t];
    } else {
        pos_lo = pos;
        t_lo = 0.0;
    }
    if (k_row < M - 1) {
        pos_hi = pos + T;
    } else if (t + 1 < T) {
        pos_hi = t + 1;
    } else {
        pos_hi = pos;
        t_hi = 0.0;
    }
    const double s = p.scal[b] * p.vec[pos];
    // bq[k]: coupling of channel l0+k with the channel below it; bq[CL]: of the chunk's last channel with the next chunk
    double bq[CL + 1];
#pragma unroll
    for (int k = 0; k <= CL; ++k) {
        const int l = l0 + k;
        bq[k] = (l >= 1 && l < L) ? s * p.cl[l - 1] : 0.0;
    }
    // Moebius chunk matrix of w -> 1 / (1 + a w):  [[0, 1], [a, 1]] per channel, later channels on the left
    double w_in;
    {
        double A = 1.0, B = 0.0, C = 0.0, D = 1.0;
#pragma unroll
        for (int k = 0; k < CL; ++k) {
            if (l0 + k < L) {
                const double a = bq[k] * bq[k];
                const double nA = C, nB = D;
                C = fma(a, A, C);
                D = fma(a, B, D);
                A = nA;
                B = nB;
            }
        }
        w_in = adi_moebius_inflow(A, B, C, D, sm, p.PW, tid, NT);  // irrelevant for chunk 0 (its first coupling is zero)
    }
    double w[CL];
    {
        double wp = w_in;
#pragma unroll
        for (int k = 0; k < CL; ++k) {
            if (l0 + k < L) wp = 1.0 / fma(bq[k] * bq[k], wp, 1.0);
            else wp = 1.0;
            w[k] = wp;
        }
    }

    // ---- load, (1 - i tau H0)_r on the fly ------------------------------------------------------
    pdl_wait();
    cplx g1[CL];
#pragma unroll
    for (int k = 0; k < CL; ++k) {
        const int l = l0 + k;
        g1[k] = c_zero();
        if (l < L) {
            const cplx *ch = p.psi + ((size_t)b * L + l) * Rp;
            const cplx g = ch[pos], glo = ch[pos_lo], ghi = ch[pos_hi];
            const cplx d = p.thd[(size_t)l * Rp + pos];
            cplx z = c_mul(d, g);
            z = c_make(fma(t_lo, glo.x, z.x), fma(t_lo, glo.y, z.y));
            z = c_make(fma(t_hi, ghi.x, z.x), fma(t_hi, ghi.y, z.y));
            g1[k] = c_make(g.x + z.y, g.y - z.x);  // g - i z
        }
    }

    // ---- forward substitution: y_l = g_l - i (beta_{l-1} w_{l-1}) y_{l-1} --------------------------
    double ef[CL];  // beta_{l-1} w_{l-1}
    ef[0] = bq[0] * w_in;
#pragma unroll
    for (int k = 1; k < CL; ++k) ef[k] = bq[k] * w[k - 1];
    cplx yin;
    {
        cplx z = g1[0], m = c_make(0.0, -ef[0]);
#pragma unroll
        for (int k = 1; k < CL; ++k) {
            z = fma_mi(ef[k], z, g1[k]);
            m = mul_mi(ef[k], m);
        }
        yin = adi_affine_inflow<true>(m, z, sm, p.PW, tid, NT);
    }
    cplx y[CL];
    y[0] = fma_mi(ef[0], yin, g1[0]);
#pragma unroll
    for (int k = 1; k < CL; ++k) y[k] = fma_mi(ef[k], y[k - 1], g1[k]);

    // ---- backward substitution: x_l = w_l y_l - i (beta_l w_l) x_{l+1} -----------------------------
    double eb[CL];
#pragma unroll
    for (int k = 0; k < CL; ++k) eb[k] = bq[k + 1] * w[k];
    cplx xin;
    {
        cplx z = c_scale(y[CL - 1], w[CL - 1]), m = c_make(0.0, -eb[CL - 1]);
#pragma unroll
        for (int k = CL - 2; k >= 0; --k) {
            z = fma_mi(eb[k], z, c_scale(y[k], w[k]));
            m = mul_mi(eb[k], m);
        }
        xin = adi_affine_inflow<false>(m, z, sm, p.PW, tid, NT);
    }
    // true inflow; out = 2 x - g1 = (1 - i tau H_int) x
    cplx x = xin;
#pragma unroll
    for (int k = CL - 1; k >= 0; --k) {
        x = fma_mi(eb[k], x, c_scale(y[k], w[k]));
        const int l = l0 + k;
        if (l < L) p.out[((size_t)b * L + l) * Rp + pos] = c_make(fma(2.0, x.x, -g1[k].x), fma(2.0, x.y, -g1[k].y));
    }
}

// ---------------------------------------------------------------------------------------------
// The closing radial pass of an ADI step: x = (1 + i tau H0)^-1 g [* mask] on every channel, out of place.
// k_unit<PROG_CN> gives a channel to one CTA of T threads x 4 rows (500 CTAs of 512 threads on 148 SMs at 2000 x 500: 3.4 waves,
// one CTA per SM).  Here a thread holds EIGHT consecutive rows -- in the M = 4 layout rows 8p .. 8p+7 are the elements [k][2p],
// [k][2p+1], k = 0..3, i.e. 32 contiguous bytes per k, so every access is still fully coalesced and nothing is transposed --
// and the LU factors go straight into registers (each thread needs exactly its own rows').  The two affine scans cost the same
// per thread whatever the rows it holds, so their cost per point halves; T/2 <= 256 threads at <= 128 registers: two CTAs per SM.
// Solve only (no 2x - g), so u = w y overwrites g.  Recurrences as cn_channel / cn8 (kernels.cuh).
// grid = (channels, batch), block = T/2 rounded up to whole warps.
// ---------------------------------------------------------------------------------------------
struct AdiRParams {
    const cplx *psi;         // [batch][L][4][T]
    cplx *out;               // [batch][L][4][T]
    const cplx *w;           // [L][4][T]  1 / pivot
    const cplx *aggP;        // [L][T]     4-row chunk multipliers, forward
    const cplx *aggQ;        // [L][T]     backward
    const double *toff;      // [4][T]     tau * h_off[i]
    const double *toff_prev; // [T]        tau * h_off[4 t - 1]
    const double *mask;      // [4][T] or nullptr
    int L, T, short_scan;
};

__global__ void __launch_bounds__(256, 2) k_adi_r(const AdiRParams p)
{
    __shared__ __align__(16) cplx sm[128];
    const int tid = threadIdx.x, NT = blockDim.x, T = p.T;
    const int l = blockIdx.x, b = blockIdx.y;
    const bool ok = 2 * tid < T;  // T is even (a multiple of 32)
    const int c0 = 2 * tid;       // first column of the thread: rows 8 tid + 4 j + k at [k][c0 + j]
    pdl_launch_dependents();
    // ---- everything that does not depend on psi ----
    const cplx *wch = p.w + (size_t)l * 4 * T;
    cplx w[8];
    double to[8];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w[4 * j + k] = ok ? ld_c(wch + k * T + c0 + j) : c_make(1.0, 0.0);
            to[4 * j + k] = ok ? p.toff[k * T + c0 + j] : 0.0;
        }
    }
    const double to_prev = ok ? p.toff_prev[c0] : 0.0;
    const cplx wprev = (ok && tid > 0) ? ld_c(wch + 3 * T + c0 - 1) : c_zero();
    cplx Pt = c_zero(), Qt = c_zero();
    if (ok) {
        Pt = c_mul(ld_c(p.aggP + (size_t)l * T + c0), ld_c(p.aggP + (size_t)l * T + c0 + 1));
        Qt = c_mul(ld_c(p.aggQ + (size_t)l * T + c0), ld_c(p.aggQ + (size_t)l * T + c0 + 1));
    }
    pdl_wait();
    const cplx *src = p.psi + ((size_t)b * p.L + l) * 4 * T;
    cplx g[8];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g[4 * j + k] = ok ? ld_c(src + k * T + c0 + j) : c_zero();
    }
    // forward, zero inflow
    cplx z = g[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) z = c_fma(e_of(to[k - 1], w[k - 1]), z, g[k]);
    const cplx yin = affine_scan_block_exclusive<true>(Pt, z, sm, sm + 32, tid, NT, p.short_scan);
    // forward, true inflow; u = w y overwrites g
    cplx y = c_fma(e_of(to_prev, wprev), yin, g[0]);
    g[0] = c_mul(w[0], y);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        y = c_fma(e_of(to[k - 1], w[k - 1]), y, g[k]);
        g[k] = c_mul(w[k], y);
    }
    // backward, zero inflow
    z = g[7];
#pragma unroll
    for (int k = 6; k >= 0; --k) z = c_fma(e_of(to[k], w[k]), z, g[k]);
    double mk[8];  // loaded late: the registers are needed above (the loads are issued before the scan's barrier)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mk[4 * j + k] = (ok && p.mask) ? p.mask[k * T + c0 + j] : 1.0;
    }
    const cplx xin = affine_scan_block_exclusive<false>(Qt, z, sm + 64, sm + 96, tid, NT, p.short_scan);
    // backward, true inflow; mask; store
    cplx *dst = p.out + ((size_t)b * p.L + l) * 4 * T;
    cplx x = c_fma(e_of(to[7], w[7]), xin, g[7]);
    if (ok) st_c(dst + 3 * T + c0 + 1, c_scale(x, mk[7]));
#pragma unroll
    for (int k = 6; k >= 0; --k) {
        x = c_fma(e_of(to[k], w[k]), x, g[k]);
        if (ok) st_c(dst + (k & 3) * T + c0 + (k >> 2), c_scale(x, mk[k]));
    }
}

// tau * h_diag of every channel, permuted to the interleaved layout: [L][M][T]
__global__ void k_make_thd(const cplx *__restrict__ h_diag, double tau, int R, int M, int T, cplx *__restrict__ thd)
{
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= M * T) return;
    const int l = blockIdx.y;
    const int k = pos / T, t = pos % T;
    const long long i = (long long)t * M + k;
    thd[(size_t)l * M * T + pos] = (i < R) ? c_scale(h_diag[(size_t)l * R + i], tau) : c_zero();
}

}  // namespace ion
