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
// sincos for rotation angles.  Every l<->l+1 rotation and every r-pair brick needs cos/sin of its own angle
// theta = (tau * field) * c_l * v_j, and in the velocity gauge the trigonometry is more than half of all FP64 work.
// The angles are small almost everywhere (config 3: |theta| < 0.08 for every h2 brick, < 0.7 for h1 beyond the first
// hundred radial rows), so the common case must not pay for a general argument reduction:
//   |theta| <= 2^-4   short Taylor polynomials                                                   ~11 FP64 instructions
//   |theta| <= pi/4   polynomials only (fdlibm's minimax kernels, < 1 ulp)                      ~16
//   |theta| <  2^20   two-term Cody-Waite reduction with FMA (exact to ~1 ulp of the remainder)  ~26
//   otherwise         CUDA's sincos (Payne-Hanek)
// Coefficients live in constant memory so that they are FMA operands, not materialised immediates.
// ---------------------------------------------------------------------------------------------
__constant__ double kSinCoef[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03,  -1.98412698298579493134e-04,
                                   2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
__constant__ double kCosCoef[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                   -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};

// sin and cos of r, |r| <= 2^-4: Taylor, truncation error < 3e-19 relative
ION_DEVINL void sincos_tiny(double r, double *sn, double *cs)
{
    const double z = r * r;
    double ps = fma(2.75573192239858906526e-06, z, -1.98412698412698412698e-04);  // 1/9!, -1/7!
    double pc = fma(2.48015873015873015873e-05, z, -1.38888888888888888889e-03);  // 1/8!, -1/6!
    ps = fma(ps, z, 8.33333333333333333333e-03);                                  // 1/5!
    pc = fma(pc, z, 4.16666666666666666667e-02);                                  // 1/4!
    ps = fma(ps, z, -1.66666666666666666667e-01);                                 // -1/3!
    *sn = fma(r * z, ps, r);
    *cs = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// sin and cos of r, |r| <= pi/4
ION_DEVINL void sincos_kernel(double r, double *sn, double *cs)
{
    const double z = r * r;
    double ps = fma(kSinCoef[5], z, kSinCoef[4]);
    double pc = fma(kCosCoef[5], z, kCosCoef[4]);
    ps = fma(ps, z, kSinCoef[3]);
    pc = fma(pc, z, kCosCoef[3]);
    ps = fma(ps, z, kSinCoef[2]);
    pc = fma(pc, z, kCosCoef[2]);
    ps = fma(ps, z, kSinCoef[1]);
    pc = fma(pc, z, kCosCoef[1]);
    ps = fma(ps, z, kSinCoef[0]);
    pc = fma(pc, z, kCosCoef[0]);
    *sn = fma(r * z, ps, r);
    *cs = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// the rare huge-angle path, out of line (and by value: no pointer escapes that would force the results into local memory)
__device__ __noinline__ double2 sincos_huge(double theta)
{
    double2 r;
    sincos(theta, &r.x, &r.y);
    return r;
}

// reduction + kernels, |theta| < 2^20
ION_DEVINL void sincos_reduced(double theta, double *sn, double *cs)
{
    const double k = rint(theta * 0.63661977236758134308);  // theta * 2/pi
    double r = fma(-k, 1.57079632679489655800e+00, theta);   // pi/2, high part
    r = fma(-k, 6.12323399573676603587e-17, r);              // pi/2, low part
    double s, c;
    sincos_kernel(r, &s, &c);
    const int q = (int)k;
    const double ss = (q & 1) ? c : s, cc = (q & 1) ? s : c;
    *sn = (q & 2) ? -ss : ss;
    *cs = ((q + 1) & 2) ? -cc : cc;
}

// cos/sin of N angles of one thread.  The path is chosen per WARP (a vote over the converged lanes), so the N
// evaluations of a path are straight-line code the scheduler can interleave, and warps never diverge here.
// `amax`: an upper bound of |theta[k]| known to the caller (e.g. |kappa| * max|v_j| for theta_j = kappa * v_j, which
// rounds monotonically), so that the selection costs one multiplication instead of N compares.
template <int N>
ION_DEVINL void fast_sincos_n_bounded(const double (&theta)[N], double (&sn)[N], double (&cs)[N], double amax)
{
    const unsigned lanes = __activemask();
    if (__all_sync(lanes, amax <= 0.0625)) {
#pragma unroll
        for (int k = 0; k < N; ++k) sincos_tiny(theta[k], &sn[k], &cs[k]);
    } else if (__all_sync(lanes, amax <= 0.78539816339744830962)) {
#pragma unroll
        for (int k = 0; k < N; ++k) sincos_kernel(theta[k], &sn[k], &cs[k]);
    } else if (__all_sync(lanes, amax < 1048576.0)) {
#pragma unroll
        for (int k = 0; k < N; ++k) sincos_reduced(theta[k], &sn[k], &cs[k]);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {  // unrolled: a dynamic index would push the arrays into local memory
            const double2 r = sincos_huge(theta[k]);
            sn[k] = r.x;
            cs[k] = r.y;
        }
    }
}
template <int N>
ION_DEVINL void fast_sincos_n(const double (&theta)[N], double (&sn)[N], double (&cs)[N])
{
    double amax = fabs(theta[0]);
#pragma unroll
    for (int k = 1; k < N; ++k) amax = fmax(amax, fabs(theta[k]));
    fast_sincos_n_bounded<N>(theta, sn, cs, amax);
}

// Programmatic dependent launch (PDL).  Every kernel of a time step depends on the wavefunction written by the
// previous one, but its prologue -- coefficient / LU-factor loads and all the trigonometry -- does not.  Kernels are
// launched with the programmatic-stream-serialization attribute: pdl_launch_dependents() at the top lets the next
// kernel's CTAs start as soon as every CTA of this one has started (i.e. during its last, partially filled wave), and
// pdl_wait() blocks until the previous grid has completed and flushed its stores; it is placed just before the first
// load of psi.  Both are no-ops for a kernel launched without the attribute.
ION_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
ION_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 16-byte asynchronous global -> shared copies (LDGSTS), used to stage the LU factors during the prologue
ION_DEVINL void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
ION_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ION_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Bulk asynchronous copies (the TMA engine's 1-D form, cp.async.bulk) signalling an mbarrier: one thread moves a whole contiguous
// block global -> shared; the consumers wait on the barrier's phase.  Used by the ensemble kernel's psi prefetch (ensemble.cuh).
ION_DEVINL unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
ION_DEVINL void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
ION_DEVINL void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
ION_DEVINL void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
ION_DEVINL void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned a = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "ION_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n"
        "@q bra ION_MBAR_DONE;\n"
        "bra ION_MBAR_WAIT;\n"
        "ION_MBAR_DONE:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}

// 128-bit global accesses of one complex128
ION_DEVINL cplx ld_c(const cplx *p) { return *p; }
ION_DEVINL void st_c(cplx *p, cplx v) { *p = v; }

// ---------------------------------------------------------------------------------------------
// l-block shards: flags and spin-waits of the peer-memory halo exchange (halo.cuh: the stand-alone exchange kernel;
// kernels.cuh: the exchange fused into the folded length-gauge step, PROG_LEN_STEP_HALO)
// ---------------------------------------------------------------------------------------------
// halo block of a shard: HF_COUNT flags (unsigned long long), then staging[side][slot][n] complex values for the exchange kernel,
// then fstage[side][slot][n] for the fused exchange (separate slots: the two mechanisms interleave at chunk boundaries)
namespace ion {
enum : int { HF_ARRIVE = 2, HF_SEQ = 4, HF_DONE = 5, HF_DONE_SIDE = 6, HF_ABORT = 8, HF_FARRIVE = 10, HF_COUNT = 16 };
}
using namespace ion;

ION_DEVINL void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ION_DEVINL void red_release_sys_add(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ION_DEVINL unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// thread 0 of a CTA waits for *p >= k; returns false on time-out / abort
ION_DEVINL bool halo_spin(const unsigned long long *p, unsigned long long k, unsigned long long *flags, long long limit)
{
    const long long t0 = clock64();
    unsigned spins = 0;
    while (ld_acquire_sys(p) < k) {
        if ((++spins & 63u) == 0u) {
            if (ld_acquire_sys(flags + HF_ABORT) != 0ull) return false;
            if (clock64() - t0 > limit) {
                st_release_sys(flags + HF_ABORT, 1ull);
                return false;
            }
        }
    }
    return true;
}

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
