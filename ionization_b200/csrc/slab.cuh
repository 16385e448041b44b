// ionization_b200 -- the velocity-gauge INTER-SOLVE kernel (sm_100a).
//
// Between the Crank-Nicolson solves of two consecutive velocity-gauge time steps the reference applies
// (evolution_methods.py:89-123 with mesh_operators.py:1204-1408)
//
//      step n:    h2_eo h2_ee   h1_o   h1_e   mask          step n+1:   h1_e   h1_o   h2_ee h2_eo
//
// i.e. five operators that alternate between the even l-pairs (2p, 2p+1) and the odd l-pairs (2p+1, 2p+2).  A kernel
// that owns l-pairs (kernels.cuh) needs one pass over psi per operator -- five passes, each bound by L2/HBM bandwidth.
// None of the five couples radial rows except the h2 bricks, which couple a row with its direct neighbours only.  So
// this kernel tiles the mesh the other way: a CTA owns a SLAB of 4G - 4 consecutive radial rows (plus two halo rows
// on either side, recomputed redundantly) of a CHUNK of l-quads (plus one halo quad on either side), applies all five
// operators with psi held in registers, and writes the result to a second buffer (out of place: the halo of one CTA
// is the interior of another).  One pass instead of five.
//
//   thread (q, g): channels 4q .. 4q+3 ("quad"), rows a + 4g .. a + 4g + 3 with a = slab * (4G - 4) - 3  (a is odd)
//      stage 1  h2 on the even l-pairs, r-sublayers in reverse order (odd, even)         scalar s_a = tau A_n
//      stage 2  h1 rotation of the odd l-pairs                                            s_a
//      stage 3  h1 rotation of the even l-pairs by s_a + s_b, radial mask
//      stage 4  h1 rotation of the odd l-pairs                                            s_b = tau A_{n+1}
//      stage 5  h2 on the even l-pairs, r-sublayers (even, odd)                           s_b
//   Both even l-pairs of a quad and the odd pair (4q+1, 4q+2) are thread-local.  The odd pairs (4q-1, 4q) and
//   (4q+3, 4q+4) straddle two threads: before stages 2 and 4 every thread publishes its channels 4q and 4q+3 in shared
//   memory and each of the two threads evaluates its own half of the rotation.
//   Because a is odd, the r-pairs (odd row, next row) of an h2 sublayer are thread-local and so is the middle pair of
//   the even sublayer; the even-sublayer pairs that straddle two row groups are completed with warp shuffles (the G
//   row groups of a quad are G consecutive lanes, G a power of two <= 32).
//   Halo: the stage order (odd, even, .., even, odd) makes the light cone two rows wide on either side in r and one
//   quad (four channels) wide on either side in l.
#pragma once
#include "kernels.cuh"

namespace ion {

struct SlabParams {
    const cplx *psi_in;    // [batch][L][4][T]
    cplx *psi_out;         // [batch][L][4][T]
    const double *vec;     // [4][T]   h1 coupling y_j, 0 in the padding
    const double *zvec;    // [4][T]   h2 r-pair coupling z_j of the pair (j, j+1), 0 for j >= R-1
    const double *mask;    // [4][T] or nullptr
    const double *cl;      // [L-1]    h1 l-pair coefficient
    const double *cl2;     // [L-1]    h2 l-pair coefficient
    const double *scal_a;  // [batch]  tau * A of the step whose tail this is
    const double *scal_b;  // [batch]  tau * A of the next step
    int L, T, R;
    int G;                 // row groups per slab; 4G rows are loaded, the middle 4G - 4 are stored
    int n_slabs;           // blockIdx.x % n_slabs
    int Qc;                // interior quads per chunk; chunk = blockIdx.x / n_slabs
    int nQ;                // quads in all: ceil(L / 4)
};

// Fused observation (north_star 4: reductions fused into the step): the state AFTER the mask of the step whose tail this
// kernel applies exists only here, between the two halves of stage 3.  k_slab<true> splits stage 3 into the rotation by s_a,
// the mask, the reductions over the CTA's interior points, and the rotation by s_b.  Every (slab, channel) / (slab, state)
// partial sum is produced by exactly one group of G lanes (shuffle reduction) and written once; k_slab_obs_reduce adds the
// slabs in a fixed order (deterministic, no atomics) into the buffers k_observe_finish assembles the record from.
// Point-wise observables only (norm, norm by l, <r>, norm within radii, inner products); <z> and <H0> couple neighbouring
// channels / rows across CTA edges and take the unfused path.
struct SlabObs {
    const double *rvec;       // [4][T]  permuted r_j
    const cplx *state_rows;   // [n_states][4][T] permuted
    const int *state_first;   // [L + 1] CSR: states of channel l
    const int *state_order;   // [n_states]
    double *partial;          // [batch][n_slabs][L][np], np = 4 + n_radii (k_observe's layout: norm_l, r_l, z_l = 0, h0_l = 0, within..)
    double *ip;               // [batch][n_slabs][n_states][2]
    double radii[ION_MAX_RADII];
    int n_radii, n_states;
    unsigned what;
    cplx *psi_n;              // STORE: [batch][L][4][T] -- the observed state itself is written out (k_slab<true, true>)
};

// position of row r (>= 0) in the row-interleaved layout with M = 4
ION_DEVINL int slab_pos(int r, int T) { return (r & 3) * T + (r >> 2); }

struct Trig {
    double c, s;
};

// r-pair brick on (lo, hi) of the Hadamard-transformed pair (S, D): S by +theta, D by -theta  (mesh_operators.py:1247-1408)
ION_DEVINL void brick(cplx &Slo, cplx &Shi, cplx &Dlo, cplx &Dhi, const Trig &t)
{
    const cplx s0 = Slo, s1 = Shi, d0 = Dlo, d1 = Dhi;
    Slo = c_make(fma(t.c, s0.x, t.s * s1.x), fma(t.c, s0.y, t.s * s1.y));
    Shi = c_make(fma(t.c, s1.x, -t.s * s0.x), fma(t.c, s1.y, -t.s * s0.y));
    Dlo = c_make(fma(t.c, d0.x, -t.s * d1.x), fma(t.c, d0.y, -t.s * d1.y));
    Dhi = c_make(fma(t.c, d1.x, t.s * d0.x), fma(t.c, d1.y, t.s * d0.y));
}
// the lower / upper member only (the partner row lives in the neighbouring lane)
ION_DEVINL void brick_lower(cplx &Slo, cplx &Dlo, const cplx Shi, const cplx Dhi, const Trig &t)
{
    Slo = c_make(fma(t.c, Slo.x, t.s * Shi.x), fma(t.c, Slo.y, t.s * Shi.y));
    Dlo = c_make(fma(t.c, Dlo.x, -t.s * Dhi.x), fma(t.c, Dlo.y, -t.s * Dhi.y));
}
ION_DEVINL void brick_upper(cplx &Shi, cplx &Dhi, const cplx Slo, const cplx Dlo, const Trig &t)
{
    Shi = c_make(fma(t.c, Shi.x, -t.s * Slo.x), fma(t.c, Shi.y, -t.s * Slo.y));
    Dhi = c_make(fma(t.c, Dhi.x, t.s * Dlo.x), fma(t.c, Dhi.y, t.s * Dlo.y));
}

// h2 on one l-pair (A = lower channel, B = upper) for the four rows of this thread.  ang[0]: pair (r0-1, r0) shared with
// the previous row group; ang[1]: (r0, r1); ang[2]: (r1, r2); ang[3]: (r2, r3); ang[4]: (r3, r3+1) shared with the next group.
// r0 is odd: ang[1], ang[3] belong to the odd sublayer, ang[0], ang[2], ang[4] to the even sublayer.
template <bool REVERSE>
ION_DEVINL void slab_h2(cplx (&A)[4], cplx (&B)[4], const Trig (&ang)[5], bool has_prev, bool has_next)
{
    // Hadamard over the l-pair without its 1/sqrt(2): the bricks are linear, both factors are applied at the end (1/2)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const cplx a = A[j], b = B[j];
        A[j] = c_add(a, b);  // sqrt(2) S
        B[j] = c_sub(a, b);  // sqrt(2) D
    }
    if (REVERSE) {
        brick(A[0], A[1], B[0], B[1], ang[1]);
        brick(A[2], A[3], B[2], B[3], ang[3]);
    }
    {   // even sublayer: (r1, r2) local; (r3, next r0) and (prev r3, r0) with the neighbouring lanes
        const cplx nS0 = shfl_down_c(A[0], 1), nD0 = shfl_down_c(B[0], 1);
        const cplx pS3 = shfl_up_c(A[3], 1), pD3 = shfl_up_c(B[3], 1);
        brick(A[1], A[2], B[1], B[2], ang[2]);
        if (has_next) brick_lower(A[3], B[3], nS0, nD0, ang[4]);
        if (has_prev) brick_upper(A[0], B[0], pS3, pD3, ang[0]);
    }
    if (!REVERSE) {
        brick(A[0], A[1], B[0], B[1], ang[1]);
        brick(A[2], A[3], B[2], B[3], ang[3]);
    }
    // back, again without the 1/sqrt(2): the result is 2 x (h2 psi); the two factors of 2 of stages 1 and 5 are folded into
    // the mask of stage 3 (every operator in between is linear; powers of two are exact)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const cplx sv = A[j], d = B[j];
        A[j] = c_add(sv, d);
        B[j] = c_sub(sv, d);
    }
}

// the five r-pair angles of one l-pair: four evaluations, the pair shared with the previous row group comes from the
// neighbouring lane (its ang[4])
ION_DEVINL void slab_h2_angles(Trig (&ang)[5], const double (&z)[5], double kappa, double zmax)
{
    double th[4], sn[4], cs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) th[j] = kappa * z[j + 1];
    fast_sincos_n_bounded<4>(th, sn, cs, fabs(kappa) * zmax);
#pragma unroll
    for (int j = 0; j < 4; ++j) ang[j + 1].c = cs[j], ang[j + 1].s = sn[j];
    ang[0].c = __shfl_up_sync(0xffffffffu, cs[3], 1);
    ang[0].s = __shfl_up_sync(0xffffffffu, sn[3], 1);
}

template <int N>
ION_DEVINL void slab_angles(Trig (&ang)[N], const double (&v)[N], double kappa, double vmax)
{
    double th[N], sn[N], cs[N];
#pragma unroll
    for (int j = 0; j < N; ++j) th[j] = kappa * v[j];
    fast_sincos_n_bounded<N>(th, sn, cs, fabs(kappa) * vmax);
#pragma unroll
    for (int j = 0; j < N; ++j) ang[j].c = cs[j], ang[j].s = sn[j];
}
// real rotation [[c, s], [-s, c]] of the l-pair (A lower, B upper)  mesh_operators.py:1204-1245
ION_DEVINL void slab_rot(cplx (&A)[4], cplx (&B)[4], const Trig (&ang)[4])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const cplx a = A[j], b = B[j];
        A[j] = c_make(fma(ang[j].c, a.x, ang[j].s * b.x), fma(ang[j].c, a.y, ang[j].s * b.y));
        B[j] = c_make(fma(ang[j].c, b.x, -ang[j].s * a.x), fma(ang[j].c, b.y, -ang[j].s * a.y));
    }
}
ION_DEVINL void slab_rot_lower(cplx (&A)[4], const cplx (&B)[4], const Trig (&ang)[4])  // A = lower member, B read-only
{
#pragma unroll
    for (int j = 0; j < 4; ++j)
        A[j] = c_make(fma(ang[j].c, A[j].x, ang[j].s * B[j].x), fma(ang[j].c, A[j].y, ang[j].s * B[j].y));
}
ION_DEVINL void slab_rot_upper(const cplx (&A)[4], cplx (&B)[4], const Trig (&ang)[4])  // B = upper member, A read-only
{
#pragma unroll
    for (int j = 0; j < 4; ++j)
        B[j] = c_make(fma(ang[j].c, B[j].x, -ang[j].s * A[j].x), fma(ang[j].c, B[j].y, -ang[j].s * A[j].y));
}

// grid = (n_slabs * n_chunks, batch), block = NT (multiple of 32, NT >= G * (Qc + 2)); dynamic smem = 12 * NT cplx
// STORE (with OBS; opt-in, ION_SLAB_OBS_STORE=1): instead of reducing in place, the kernel writes psi_n -- every interior point exactly
// once -- into a second buffer and the stand-alone k_observe reduces it on a side branch of the stream / captured graph (every observable
// k_observe knows rides along, <z> and <H0> included).  Measured slower on C3 (47.1 vs 42.7 us per observed step): k_observe's one-channel
// CTAs do not fit into the SM time the 1.7-wave pair kernel of the next step leaves idle, and the main stream ends up waiting for them.
template <bool OBS, bool STORE = false>
__global__ void __launch_bounds__(512, 1) k_slab(const SlabParams p, const SlabObs o)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *xch = reinterpret_cast<cplx *>(smem_raw);  // [2 edges + the upper straddling pair's cos/sin][4 rows][NT]

    const int tid = threadIdx.x, NT = blockDim.x;
    const int G = p.G, T = p.T, L = p.L;
    const int g = tid & (G - 1), ql = tid / G;
    const int slab = blockIdx.x % p.n_slabs, chunk = blockIdx.x / p.n_slabs;
    const int b = blockIdx.y;
    const int q_int0 = chunk * p.Qc, q_int1 = min(q_int0 + p.Qc, p.nQ);
    const int q_first = q_int0 - (chunk > 0 ? 1 : 0);
    const int q = q_first + ql;
    const bool q_ok = q < min(q_int1 + 1, p.nQ);
    const int l0 = 4 * q;
    const int W = 4 * G - 4;
    const int r0 = slab * W - 3 + 4 * g;  // first row of this thread (odd; may be negative or beyond R)

    pdl_launch_dependents();
#ifdef ION_EXP_CLOCKS
    long long ck[10];
#define ION_SCK(i) ck[i] = clock64()
#else
#define ION_SCK(i)
#endif
    ION_SCK(0);

    // ---- coefficients of this thread's rows (independent of psi) ----
    double v[4], mk[4], z[5];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + j;
        const bool ok = r >= 0 && r < p.R;
        v[j] = ok ? p.vec[slab_pos(r, T)] : 0.0;
        // 1/4: the unnormalised Hadamard pairs of stages 1 and 5.  Unobserved kernel: the mask itself is loaded where it is used (stage 3)
        if constexpr (OBS) mk[j] = 0.25 * ((ok && p.mask) ? p.mask[slab_pos(r, T)] : 1.0);
        else mk[j] = 0.25;
    }
    const double vmax = fmax(fmax(fabs(v[0]), fabs(v[1])), fmax(fabs(v[2]), fabs(v[3])));
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int r = r0 - 1 + j;  // lower row of the pair
        z[j] = (r >= 0 && r + 1 < p.R) ? p.zvec[slab_pos(r, T)] : 0.0;
    }
    const double zmax = fmax(fmax(fabs(z[1]), fabs(z[2])), fmax(fabs(z[3]), fabs(z[4])));
    const double sa = p.scal_a[b], sb = p.scal_b[b];
    auto coef = [&](const double *c, int l) -> double { return (q_ok && l >= 0 && l + 1 < L) ? c[l] : 0.0; };
    const bool has_prev = g > 0, has_next = g + 1 < G;

    // ---- load psi: 4 channels x 4 rows, interleaved with the trigonometry of stage 1 ----
    // 16 warps issuing 16 loads each saturate the SM's load path for ~2.5 k cycles (the issue itself stalls), so the angles of the
    // first pair are evaluated before the wait (they overlap the previous kernel's tail), those of the second pair between the two
    // batches of loads.
    cplx X[4][4];
    Trig ang01[5], ang23[5];
    slab_h2_angles(ang01, z, sa * coef(p.cl2, l0), zmax);
    const double kap23 = sa * coef(p.cl2, l0 + 2);
    ION_SCK(1);
    pdl_wait();
    ION_SCK(2);
    const cplx *ibase = p.psi_in + ((size_t)b * L + l0) * 4 * T;
    auto load_pair = [&](int c0) {
#pragma unroll
        for (int c = c0; c < c0 + 2; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = r0 + j;
                const bool ok = q_ok && (l0 + c < L) && r >= 0 && r < p.R;
                X[c][j] = ok ? ld_c(ibase + (size_t)c * 4 * T + slab_pos(r, T)) : c_zero();
            }
        }
    };
    load_pair(0);
    asm volatile("" ::: "memory");
    slab_h2_angles(ang23, z, kap23, zmax);
    asm volatile("" ::: "memory");
    load_pair(2);

    // ---- stage 1: h2 (reversed) on (0,1), (2,3) with s_a ----
    ION_SCK(3);
    slab_h2<true>(X[0], X[1], ang01, has_prev, has_next);
    slab_h2<true>(X[2], X[3], ang23, has_prev, has_next);
    // ---- stages 2 and 4: odd l-pairs; stage 3 in between ----
    ION_SCK(4);
    const int up = tid + G, dn = tid - G;  // threads holding the same rows of the next / previous quad
    // the two passes (stages 2 / 4) are unrolled in the unobserved kernel -- the compiler then keeps what only pass 0 needs (stage 3) out of
    // pass 1's live ranges: k_slab 18.9 -> 17.6 us per launch, C3 VEL 27.27 -> 26.47 us per step together with the late mask load below;
    // the observed variants keep the rolled loop (their stage 3 is three times the code)
#pragma unroll(OBS ? 1 : 2)
    for (int pass = 0; pass < 2; ++pass) {
        const double s = pass == 0 ? sa : sb;
        Trig ang_up[4];  // pair (l0 + 3, l0 + 4): evaluated here, handed to the next quad's thread with the edge channel
        slab_angles<4>(ang_up, v, s * coef(p.cl, l0 + 3), vmax);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            xch[(0 * 4 + j) * NT + tid] = X[0][j];
            xch[(1 * 4 + j) * NT + tid] = X[3][j];
            xch[(2 * 4 + j) * NT + tid] = c_make(ang_up[j].c, ang_up[j].s);
        }
        __syncthreads();
        {
            Trig ang[4];
            slab_angles<4>(ang, v, s * coef(p.cl, l0 + 1), vmax);
            slab_rot(X[1], X[2], ang);
        }
        if (ql > 0) {  // pair (l0 - 1, l0): my channel 0 is the upper member
            cplx nb[4];
            Trig ang[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                nb[j] = xch[(1 * 4 + j) * NT + dn];
                const cplx t = xch[(2 * 4 + j) * NT + dn];
                ang[j].c = t.x, ang[j].s = t.y;
            }
            slab_rot_upper(nb, X[0], ang);
        }
        if (up < NT) {  // pair (l0 + 3, l0 + 4): my channel 3 is the lower member
            cplx nb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) nb[j] = xch[(0 * 4 + j) * NT + up];
            slab_rot_lower(X[3], nb, ang_up);
        }
        if (pass == 0) {
            // ---- stage 3: even l-pairs by s_a + s_b, mask ----
            Trig ang[4];
            if constexpr (!OBS) {
                if (p.mask) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = r0 + j;
                        if (r >= 0 && r < p.R) mk[j] = 0.25 * p.mask[slab_pos(r, T)];
                    }
                }
                slab_angles<4>(ang, v, (sa + sb) * coef(p.cl, l0), vmax);
                slab_rot(X[0], X[1], ang);
                slab_angles<4>(ang, v, (sa + sb) * coef(p.cl, l0 + 2), vmax);
                slab_rot(X[2], X[3], ang);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) X[c][j] = c_scale(X[c][j], mk[j]);
                }
            } else {
                // tail of step n: h1_e(s_a), mask.  X holds 2 x psi (the unnormalised Hadamard pair of stage 1), so 2 mk = mask / 2
                // gives psi_n exactly (powers of two are exact); the other factor 1/2 (stage 5) follows the observation
                slab_angles<4>(ang, v, sa * coef(p.cl, l0), vmax);
                slab_rot(X[0], X[1], ang);
                slab_angles<4>(ang, v, sa * coef(p.cl, l0 + 2), vmax);
                slab_rot(X[2], X[3], ang);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) X[c][j] = c_scale(X[c][j], 2.0 * mk[j]);
                }
                if constexpr (STORE) {   // ---- psi_n, interior points only (the stage-5 store pattern) ----
                    if (q_ok && q >= q_int0 && q < q_int1) {
                        cplx *nbase = o.psi_n + ((size_t)b * L + l0) * 4 * T;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int idx = 4 * g + j, r = r0 + j;
                                if (l0 + c < L && idx >= 2 && idx < 4 * G - 2 && r >= 0 && r < p.R) st_c(nbase + (size_t)c * 4 * T + slab_pos(r, T), X[c][j]);
                            }
                        }
                    }
                } else {   // ---- reductions over this thread's interior points; the G lanes of a quad are consecutive lanes ----
                    const bool q_in = q_ok && q >= q_int0 && q < q_int1;
                    double rr[4];
                    bool in[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int idx = 4 * g + j, r = r0 + j;
                        in[j] = q_in && idx >= 2 && idx < 4 * G - 2 && r >= 0 && r < p.R;
                        rr[j] = (in[j] && o.rvec) ? o.rvec[slab_pos(r, T)] : 0.0;
                    }
                    const int np = 4 + o.n_radii;
                    const bool with_ip = (o.what & 2u) && o.n_states > 0;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        const int l = l0 + c;
                        const bool ch_ok = q_in && l < L;
                        double n2[4], a0 = 0.0, a1 = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            // X[c] with a loop-variant c would live in local memory: select from the four channels
                            const cplx x = c == 0 ? X[0][j] : (c == 1 ? X[1][j] : (c == 2 ? X[2][j] : X[3][j]));
                            n2[j] = (in[j] && ch_ok) ? c_abs2(x) : 0.0;
                            a0 += n2[j];
                            a1 += rr[j] * n2[j];
                        }
                        for (int s = G >> 1; s > 0; s >>= 1) {
                            a0 += __shfl_down_sync(0xffffffffu, a0, s);
                            a1 += __shfl_down_sync(0xffffffffu, a1, s);
                        }
                        double *out = o.partial + (((size_t)b * p.n_slabs + slab) * L + (ch_ok ? l : 0)) * np;
                        if (g == 0 && ch_ok) {
                            out[0] = a0;
                            out[1] = a1;
                            out[2] = 0.0;
                            out[3] = 0.0;
                        }
                        for (int k = 0; k < o.n_radii; ++k) {  // norm within radius k (mesh/data.py:419-422); uniform trip count
                            double w = 0.0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) w += (rr[j] <= o.radii[k]) ? n2[j] : 0.0;
                            for (int s = G >> 1; s > 0; s >>= 1) w += __shfl_down_sync(0xffffffffu, w, s);
                            if (g == 0 && ch_ok) out[4 + k] = w;
                        }
                        if (with_ip) {
                            // inner products with the test states of this channel (mesh/meshes.py:1117-1129): sum conj(row) psi.
                            // The trip count is made uniform over the warp (its four quads hold different channels).
                            const int first = ch_ok ? o.state_first[l] : 0, ns = ch_ok ? o.state_first[l + 1] - first : 0;
                            int ns_max = ns;
#pragma unroll
                            for (int s = 16; s > 0; s >>= 1) ns_max = max(ns_max, __shfl_xor_sync(0xffffffffu, ns_max, s));
                            for (int si = 0; si < ns_max; ++si) {
                                const bool have = si < ns;
                                const int st = have ? o.state_order[first + si] : 0;
                                double re = 0.0, im = 0.0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (have && in[j]) {
                                        const cplx x = c == 0 ? X[0][j] : (c == 1 ? X[1][j] : (c == 2 ? X[2][j] : X[3][j]));
                                        const cplx a = o.state_rows[(size_t)st * 4 * T + slab_pos(r0 + j, T)];
                                        re += a.x * x.x + a.y * x.y;
                                        im += a.x * x.y - a.y * x.x;
                                    }
                                }
                                for (int s = G >> 1; s > 0; s >>= 1) {
                                    re += __shfl_down_sync(0xffffffffu, re, s);
                                    im += __shfl_down_sync(0xffffffffu, im, s);
                                }
                                if (have && g == 0) {
                                    double *oi = o.ip + (((size_t)b * p.n_slabs + slab) * o.n_states + st) * 2;
                                    oi[0] = re;
                                    oi[1] = im;
                                }
                            }
                        }
                    }
                }
                // head of step n + 1: h1_e(s_b); the remaining 1/2
                slab_angles<4>(ang, v, sb * coef(p.cl, l0), vmax);
                slab_rot(X[0], X[1], ang);
                slab_angles<4>(ang, v, sb * coef(p.cl, l0 + 2), vmax);
                slab_rot(X[2], X[3], ang);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) X[c][j] = c_scale(X[c][j], 0.5);
                }
            }
            __syncthreads();  // everybody has read the stage-2 edges: the exchange buffer may be overwritten
        }
    }
    // ---- stage 5: h2 (forward) on (0,1), (2,3) with s_b; the first pair's stores are issued before the second pair is computed ----
    ION_SCK(5);
    const bool st_ok = q_ok && q >= q_int0 && q < q_int1;
    cplx *obase = p.psi_out + ((size_t)b * L + l0) * 4 * T;
    auto store_pair = [&](int c0) {  // the interior: rows with 2 <= 4g + j < 4G - 2 of the interior quads
        if (!st_ok) return;
#pragma unroll
        for (int c = c0; c < c0 + 2; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = 4 * g + j, r = r0 + j;
                if (l0 + c < L && idx >= 2 && idx < 4 * G - 2 && r >= 0 && r < p.R) st_c(obase + (size_t)c * 4 * T + slab_pos(r, T), X[c][j]);
            }
        }
    };
    {
        Trig ang[5];
        slab_h2_angles(ang, z, sb * coef(p.cl2, l0), zmax);
        slab_h2<false>(X[0], X[1], ang, has_prev, has_next);
        ION_SCK(6);
        store_pair(0);
        asm volatile("" ::: "memory");
        slab_h2_angles(ang, z, sb * coef(p.cl2, l0 + 2), zmax);
        slab_h2<false>(X[2], X[3], ang, has_prev, has_next);
        store_pair(2);
    }
#ifdef ION_EXP_CLOCKS
    ION_SCK(7);
    if ((tid == 0 || tid == 300) && (blockIdx.x == 3 || blockIdx.x == 120) && blockIdx.y == 0)
        printf("SCK cta %d t %d: coef %lld pdlwait %lld loadissue %lld stage1 %lld stages2-4 %lld stage5 %lld store %lld total %lld\n", (int)blockIdx.x, tid,
               ck[1] - ck[0], ck[2] - ck[1], ck[3] - ck[2], ck[4] - ck[3], ck[5] - ck[4], ck[6] - ck[5], ck[7] - ck[6], ck[7] - ck[0]);
#endif
}

// Record assembly for k_slab<true>: (1) one CTA per channel (and one per 64 inner-product components) adds the per-slab partial
// sums -- thread = slab, combined by block_sum's fixed tree -> partial [batch][L][np], ip_out [batch][n_states][2] (x ipm),
// i.e. exactly what k_observe leaves behind; (2) the CTA that finishes last for a simulation (a counter per simulation)
// assembles the record the way k_observe_finish does.  Fixed summation orders everywhere: deterministic, no floating-point
// atomics.  Runs on a side branch of the captured graph (engine.cu: launch_slab).  grid = (L + ceil(2 n_states / 64), batch), block = 128.
__global__ void __launch_bounds__(128) k_slab_obs_assemble(const double *__restrict__ sp, const double *__restrict__ sip, double *__restrict__ partial,
                                                           double *__restrict__ ip_out, unsigned *__restrict__ counter, double *__restrict__ out, int n_slabs, int L,
                                                           int n_states, int n_radii, unsigned what, double ipm, long long rec)
{
    __shared__ double sm[32 * (4 + ION_MAX_RADII)];
    __shared__ unsigned last_sm;
    const int tid = threadIdx.x, b = blockIdx.y;
    const int np = 4 + n_radii, n1 = L * np, n2 = 2 * n_states;
    if ((int)blockIdx.x < L) {
        const int l = blockIdx.x;
        double acc[4 + ION_MAX_RADII];
#pragma unroll
        for (int q = 0; q < 4 + ION_MAX_RADII; ++q) acc[q] = 0.0;
        for (int s = tid; s < n_slabs; s += blockDim.x) {
            const double *src = sp + (((size_t)b * n_slabs + s) * L + l) * np;
#pragma unroll
            for (int q = 0; q < 4 + ION_MAX_RADII; ++q)
                if (q < np) acc[q] += src[q];
        }
        block_sum<4 + ION_MAX_RADII>(acc, sm, tid, blockDim.x);
        if (tid == 0) {
#pragma unroll
            for (int q = 0; q < 4 + ION_MAX_RADII; ++q)
                if (q < np) partial[(size_t)b * n1 + (size_t)l * np + q] = acc[q];
        }
    } else {
        // 64 components per CTA, two threads (even / odd slabs) per component, combined in a fixed order
        const int j = ((int)blockIdx.x - L) * 64 + (tid >> 1), half = tid & 1;
        double acc = 0.0;
        if (j < n2)
            for (int s = half; s < n_slabs; s += 2) acc += sip[((size_t)b * n_slabs + s) * n2 + j];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (j < n2 && half == 0) ip_out[(size_t)b * n2 + j] = acc * ipm;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last_sm = (atomicAdd(counter + b, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!last_sm) return;
    if (tid == 0) counter[b] = 0u;  // ready for the next observation
    __threadfence();
    const volatile double *pb = partial + (size_t)b * n1;
    const volatile double *ipb = ip_out + (size_t)b * n2;
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
        for (int s = tid; s < n2; s += blockDim.x) o[c + s] = ipb[s];
        c += n2;
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

}  // namespace ion