// ionization_b200 -- shared device helpers (sm_100a, FP64 complex arithmetic, warp/CTA affine scans)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;  // (re, im) -- same memory layout as numpy complex128 / C99 double complex

#define ION_DEVINL __device__ __forceinline__

ION_DEVINL cplx c_make(double re, double im) { return make_double2(re, im); }
ION_DEVINL cplx c_zero() { return make_double2(0.0, 0.0); }
ION_DEVINL cplx c_add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
ION_DEVINL cplx c_sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
ION_DEVINL cplx c_scale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
ION_DEVINL cplx c_mul(cplx a, cplx b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
// a*b + c
ION_DEVINL cplx c_fma(cplx a, cplx b, cplx c)
{
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
ION_DEVINL cplx c_conj(cplx a) { return make_double2(a.x, -a.y); }
ION_DEVINL double c_abs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }
// 1 / a  (Smith-free: |a| is O(1) for Crank-Nicolson pivots; plain formula, as numpy's would round)
ION_DEVINL cplx c_inv(cplx a)
{
    double d = 1.0 / c_abs2(a);
    return make_double2(a.x * d, -a.y * d);
}

ION_DEVINL cplx shfl_up_c(cplx v, int d)
{
    return make_double2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
ION_DEVINL cplx shfl_down_c(cplx v, int d)
{
    return make_double2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
ION_DEVINL cplx shfl_c(cplx v, int src)
{
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// ---------------------------------------------------------------------------------------------
// Incremental sincos.  A thread needs cos/sin of several nearby angles theta_k = sc * vec[k] (consecutive radial
// rows: the angle varies smoothly with r).  One full-range sincos is taken at a base angle; the others are
// obtained by rotating the base by d = theta_k - theta_0 with a Taylor polynomial when |d| < 2^-6 (truncation
// error < 2e-22 relative, i.e. below double rounding) and by a full sincos otherwise.  Always derived from the
// base, never chained, so errors do not accumulate.
// ---------------------------------------------------------------------------------------------
struct SinCosBase {
    double theta, c, s;
};
ION_DEVINL SinCosBase sincos_base(double theta)
{
    SinCosBase b;
    b.theta = theta;
    sincos(theta, &b.s, &b.c);
    return b;
}
ION_DEVINL void sincos_near(const SinCosBase &b, double theta, double *sn, double *cs)
{
    const double d = theta - b.theta;
    if (fabs(d) < 0.015625) {
        const double d2 = d * d;
        // sin d = d (1 - d2/6 (1 - d2/20 (1 - d2/42)));  1 - cos d = d2/2 (1 - d2/12 (1 - d2/30 (1 - d2/56)))
        const double sd = d * fma(-d2 * (1.0 / 6.0), fma(-d2 * (1.0 / 20.0), fma(-d2, 1.0 / 42.0, 1.0), 1.0), 1.0);
        const double q = 0.5 * d2 * fma(-d2 * (1.0 / 12.0), fma(-d2 * (1.0 / 30.0), fma(-d2, 1.0 / 56.0, 1.0), 1.0), 1.0);
        *cs = b.c - fma(b.c, q, b.s * sd);
        *sn = b.s - fma(b.s, q, -b.c * sd);
    } else {
        sincos(theta, sn, cs);
    }
}

// Programmatic dependent launch (PDL).  Every kernel of a time step depends on the wavefunction written by the
// previous one, but its prologue -- coefficient / LU-factor loads and all the trigonometry -- does not.  Kernels are
// launched with the programmatic-stream-serialization attribute: pdl_launch_dependents() at the top lets the next
// kernel's CTAs start as soon as every CTA of this one has started (i.e. during its last, partially filled wave), and
// pdl_wait() blocks until the previous grid has completed and flushed its stores; it is placed just before the first
// load of psi.  Both are no-ops for a kernel launched without the attribute.
ION_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
ION_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 128-bit global accesses of one complex128
ION_DEVINL cplx ld_c(const cplx *p) { return *p; }
ION_DEVINL void st_c(cplx *p, cplx v) { *p = v; }

// ---------------------------------------------------------------------------------------------
// Affine-map scans.  Thread t carries the map f_t(v) = P_t * v + B_t.
// FWD:  value entering thread t is (f_{t-1} o ... o f_0)(0)
// !FWD: value entering thread t is (f_{t+1} o ... o f_{T-1})(0)        (reverse direction)
// Kogge-Stone over the 32 lanes with shuffles, then over the <= 32 warp aggregates through
// shared memory (every warp redoes the tiny second-level scan, so one __syncthreads suffices).
// ---------------------------------------------------------------------------------------------
template <bool FWD>
ION_DEVINL void affine_scan_warp(cplx &P, cplx &B, int lane)
{
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        cplx Pp = FWD ? shfl_up_c(P, s) : shfl_down_c(P, s);
        cplx Bp = FWD ? shfl_up_c(B, s) : shfl_down_c(B, s);
        bool act = FWD ? (lane >= s) : (lane + s < 32);
        if (act) {
            B = c_fma(P, Bp, B);
            P = c_mul(P, Pp);
        }
    }
}

// smP/smB: 32 entries each, private to this call site (no reuse hazard inside one kernel phase).
// `reach` > 0: the caller guarantees (k_scan_bound, checked on the host when the LU factors are built) that the
// product of the multipliers over any `reach` whole warps is below 1e-30 in magnitude, so the inflow of a warp is a
// Horner sum over its reach+1 nearest neighbours; everything dropped is < 1e-30 relative to the largest entry times a
// further full-warp product.  reach == 0: full block-wide scan.
template <bool FWD>
ION_DEVINL cplx affine_scan_block_exclusive(cplx P, cplx B, cplx *smP, cplx *smB, int tid, int nthreads, int reach)
{
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
    affine_scan_warp<FWD>(P, B, lane);
    cplx win = c_zero();
    if (nw > 1) {
        if (lane == (FWD ? 31 : 0)) {
            smP[warp] = P;
            smB[warp] = B;
        }
        __syncthreads();
        if (reach == 1) {  // the common case, kept branch-light
            const int src = FWD ? warp - 1 : warp + 1, src2 = FWD ? warp - 2 : warp + 2;
            if (src >= 0 && src < nw) {
                win = smB[src];
                if (src2 >= 0 && src2 < nw) win = c_fma(smP[src], smB[src2], win);
            }
        } else if (reach > 1) {
#pragma unroll 1
            for (int j = reach + 1; j >= 1; --j) {
                const int src = FWD ? warp - j : warp + j;
                if (src >= 0 && src < nw) win = c_fma(smP[src], win, smB[src]);
            }
        } else {
            cplx wP = (lane < nw) ? smP[lane] : c_make(1.0, 0.0);
            cplx wB = (lane < nw) ? smB[lane] : c_zero();
            affine_scan_warp<FWD>(wP, wB, lane);
            int src = FWD ? warp - 1 : warp + 1;
            cplx v = shfl_c(wB, src & 31);
            win = (src >= 0 && src < nw) ? v : c_zero();
        }
    }
    // exclusive inside the warp
    cplx Pe = FWD ? shfl_up_c(P, 1) : shfl_down_c(P, 1);
    cplx Be = FWD ? shfl_up_c(B, 1) : shfl_down_c(B, 1);
    bool first = FWD ? (lane == 0) : (lane == 31);
    return first ? win : c_fma(Pe, win, Be);
}

// ---------------------------------------------------------------------------------------------
// 2x2 complex matrices acting projectively on a pivot p = n/d: the Thomas pivot recurrence
// p_i = D_i + o^2 / p_{i-1} is the Moebius map [[D_i, o^2], [1, 0]].  Products are rescaled by a power of two
// (pivots have modulus > 1, so unnormalised products overflow after a few hundred rows).
// ---------------------------------------------------------------------------------------------
struct Mat2 {
    cplx a, b, c, d;
};
ION_DEVINL Mat2 mat_mul(const Mat2 &L, const Mat2 &R)  // L * R  (R acts first)
{
    Mat2 o;
    o.a = c_fma(L.b, R.c, c_mul(L.a, R.a));
    o.b = c_fma(L.b, R.d, c_mul(L.a, R.b));
    o.c = c_fma(L.d, R.c, c_mul(L.c, R.a));
    o.d = c_fma(L.d, R.d, c_mul(L.c, R.b));
    return o;
}
ION_DEVINL void mat_normalize(Mat2 &m)
{
    double mx = fmax(fmax(fmax(fabs(m.a.x), fabs(m.a.y)), fmax(fabs(m.b.x), fabs(m.b.y))),
                     fmax(fmax(fabs(m.c.x), fabs(m.c.y)), fmax(fabs(m.d.x), fabs(m.d.y))));
    int e = ((__double2hiint(mx) >> 20) & 0x7ff) - 1023;      // floor(log2(mx)) for normal numbers
    e = max(-1000, min(1000, e));
    const double sc = __hiloint2double((1023 - e) << 20, 0);  // 2^-e, exact
    m.a = c_scale(m.a, sc);
    m.b = c_scale(m.b, sc);
    m.c = c_scale(m.c, sc);
    m.d = c_scale(m.d, sc);
}
ION_DEVINL Mat2 mat_shfl_up(const Mat2 &m, int d)
{
    Mat2 o;
    o.a = shfl_up_c(m.a, d);
    o.b = shfl_up_c(m.b, d);
    o.c = shfl_up_c(m.c, d);
    o.d = shfl_up_c(m.d, d);
    return o;
}
ION_DEVINL cplx c_div(cplx n, cplx d)
{
    const double s = 1.0 / c_abs2(d);
    return make_double2((n.x * d.x + n.y * d.y) * s, (n.y * d.x - n.x * d.y) * s);
}

// block-wide sum of NV doubles; result valid in thread 0.  sm: >= 32*NV doubles.
template <int NV>
ION_DEVINL void block_sum(double (&v)[NV], double *sm, int tid, int nthreads)
{
    const int lane = tid & 31, warp = tid >> 5, nw = (nthreads + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], s);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sm[warp * NV + i] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < nw) ? sm[lane * NV + i] : 0.0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) x += __shfl_down_sync(0xffffffffu, x, s);
            v[i] = x;
        }
    }
    __syncthreads();
}
