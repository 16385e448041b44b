/* ORACLE TEST INFRASTRUCTURE -- not product code.
 *
 * Plain-C restatement of the reference's mesh time-evolution hot path, used
 * (1) as a fast checker at sizes the numpy oracle is too slow for and (2) as
 * the CPU baseline ("port") timed by bench.py.  It follows the reference's
 * operator ORDER and arithmetic (explicit tridiagonal mat-vec, then the Thomas
 * recurrence of cy.pyx:28-48 with its two divisions per row); it does not use
 * any of the algebraic shortcuts of the CUDA engine.  Checked against
 * oracle/restate.py and tests/golden in tests/test_oracle_pinned.py.
 *
 * Citations are relative to /root/reference/ionization.
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex c128;

/* ---- a1: Thomas algorithm, cy.pyx:9-50 ------------------------------------ */
/* sub[i] couples row i+1 to x[i]; sup[i] couples row i to x[i+1]; length n-1. */
void ora_tdma(long n, const c128 *sub, const c128 *diag, const c128 *sup, const c128 *d, c128 *x, c128 *work /* 2n */)
{
    c128 *cp = work, *dp = work + n;
    if (n == 1) { x[0] = d[0] / diag[0]; return; }
    cp[0] = sup[0] / diag[0];            /* cy.pyx:29 */
    dp[0] = d[0] / diag[0];              /* cy.pyx:30 */
    for (long i = 1; i < n - 1; ++i) {   /* cy.pyx:31-38 */
        c128 s = sub[i - 1];
        c128 denom = diag[i] - s * cp[i - 1];
        cp[i] = sup[i] / denom;
        dp[i] = (d[i] - s * dp[i - 1]) / denom;
    }
    dp[n - 1] = (d[n - 1] - sub[n - 2] * dp[n - 2]) / (diag[n - 1] - sub[n - 2] * cp[n - 2]); /* :40 */
    x[n - 1] = dp[n - 1];                /* cy.pyx:43 */
    for (long i = n - 2; i >= 0; --i)    /* cy.pyx:44-45 */
        x[i] = dp[i] - cp[i] * x[i + 1];
}

/* ---- Crank-Nicolson in r for one channel: evolution_methods.py:98-111 ------ */
static void cn_row(long R, c128 *g, const c128 *hd, const double *off, double tau, c128 *work /* 6R */)
{
    c128 *rhs = work, *sub = work + R, *dia = work + 2 * R, *tw = work + 3 * R; /* tw: 2R */
    c128 *x = work + 5 * R;
    for (long j = 0; j < R; ++j) {
        c128 v = (1.0 - I * tau * hd[j]) * g[j];            /* DotOperator, mesh_operators.py:104-106 */
        if (j > 0) v += (-I * tau * off[j - 1]) * g[j - 1];
        if (j < R - 1) v += (-I * tau * off[j]) * g[j + 1];
        rhs[j] = v;
        dia[j] = 1.0 + I * tau * hd[j];
        if (j < R - 1) sub[j] = I * tau * off[j];
    }
    ora_tdma(R, sub, dia, sub, rhs, x, tw);                  /* TDMAOperator, :109-111 */
    memcpy(g, x, (size_t)R * sizeof(c128));
}

/* parity of the flat F-order index of the lower pair member: mesh_operators.py:1045 */
static inline int flat_parity(long L, long l, long j) { return (int)((j * L + l) & 1); }

/* ---- a7: one length-gauge sweep, mesh_operators.py:1037-1080 ---------------- */
static void len_sweep(long L, long R, c128 *g, const double *c_l, const double *x_j, double s, int parity)
{
#pragma omp parallel for schedule(static)
    for (long j = 0; j < R; ++j) {
        for (long l = 0; l < L - 1; ++l) {
            if (flat_parity(L, l, j) != parity) continue;
            double a = s * c_l[l] * x_j[j];
            double c = cos(a), sn = sin(a);
            c128 lo = g[l * R + j], hi = g[(l + 1) * R + j];
            g[l * R + j] = c * lo - I * sn * hi;
            g[(l + 1) * R + j] = -I * sn * lo + c * hi;
        }
    }
}

static void cn_all(long L, long R, c128 *g, const c128 *h_diag, const double *h_off, double tau)
{
#pragma omp parallel
    {
        c128 *work = (c128 *)malloc((size_t)(6 * R) * sizeof(c128));
#pragma omp for schedule(static)
        for (long l = 0; l < L; ++l) cn_row(R, g + l * R, h_diag + l * R, h_off, tau, work);
        free(work);
    }
}

static void apply_mask(long L, long R, c128 *g, const double *mask)
{
#pragma omp parallel for schedule(static)
    for (long l = 0; l < L; ++l)
        for (long j = 0; j < R; ++j) g[l * R + j] *= mask[j];     /* mesh/meshes.py:257 */
}

/* SplitInteractionOperator, length gauge: evolution_methods.py:89-123 */
void ora_sh_len_so_steps(long L, long R, c128 *g, const c128 *h_diag, const double *h_off, const double *c_l,
                         const double *x_j, const double *mask, long nsteps, const double *taus, const double *fields)
{
    for (long n = 0; n < nsteps; ++n) {
        double tau = taus[n], s = tau * fields[n];
        len_sweep(L, R, g, c_l, x_j, s, 0);
        len_sweep(L, R, g, c_l, x_j, s, 1);
        cn_all(L, R, g, h_diag, h_off, tau);
        len_sweep(L, R, g, c_l, x_j, s, 1);
        len_sweep(L, R, g, c_l, x_j, s, 0);
        apply_mask(L, R, g, mask);
    }
}

/* ---- a8: velocity gauge, mesh_operators.py:1204-1408, :150-204 -------------- */
static void h1_sweep(long L, long R, c128 *g, const double *f1_l, const double *y_j, double s, int parity)
{
#pragma omp parallel for schedule(static)
    for (long j = 0; j < R; ++j) {
        for (long l = 0; l < L - 1; ++l) {
            if (flat_parity(L, l, j) != parity) continue;
            double a = s * f1_l[l] * y_j[j];
            double c = cos(a), sn = sin(a);
            c128 lo = g[l * R + j], hi = g[(l + 1) * R + j];
            g[l * R + j] = c * lo + sn * hi;                  /* split_h1 :1204-1245 */
            g[(l + 1) * R + j] = -sn * lo + c * hi;
        }
    }
}

static void h2_sweep(long L, long R, c128 *g, const double *c_l, const double *z_j, double s, int lpar, int rpar)
{
    const double rs2 = 1.0 / sqrt(2.0);
#pragma omp parallel for schedule(static)
    for (long l = lpar; l < L - 1; l += 2) {                  /* SimilarityOperator :150-204 */
        c128 *a = g + l * R, *b = g + (l + 1) * R;
        for (long j = rpar; j < R - 1; j += 2) {
            double th = s * c_l[l] * z_j[j];
            double c = cos(th), sn = sin(th);
            c128 s0 = (a[j] + b[j]) * rs2, d0 = (a[j] - b[j]) * rs2;
            c128 s1 = (a[j + 1] + b[j + 1]) * rs2, d1 = (a[j + 1] - b[j + 1]) * rs2;
            c128 ns0 = c * s0 + sn * s1, ns1 = -sn * s0 + c * s1;   /* split_h2 :1247-1408 */
            c128 nd0 = c * d0 - sn * d1, nd1 = sn * d0 + c * d1;
            a[j] = (ns0 + nd0) * rs2; b[j] = (ns0 - nd0) * rs2;
            a[j + 1] = (ns1 + nd1) * rs2; b[j + 1] = (ns1 - nd1) * rs2;
        }
        /* points not in any r-pair of this parity: Hadamard twice = identity up to rounding;
           the reference multiplies by 1/sqrt2 twice -- reproduce that rounding path */
        for (long j = 0; j < R; ++j) {
            int paired = (j >= rpar) && (j < rpar + 2 * ((R - 1 - rpar + 1) / 2));
            if (paired) continue;
            c128 s0 = (a[j] + b[j]) * rs2, d0 = (a[j] - b[j]) * rs2;
            a[j] = (s0 + d0) * rs2; b[j] = (s0 - d0) * rs2;
        }
    }
}

void ora_sh_vel_so_steps(long L, long R, c128 *g, const c128 *h_diag, const double *h_off, const double *c_l,
                         const double *f1_l, const double *y_j, const double *z_j, const double *mask, long nsteps,
                         const double *taus, const double *fields)
{
    for (long n = 0; n < nsteps; ++n) {
        double tau = taus[n], s = tau * fields[n];
        h1_sweep(L, R, g, f1_l, y_j, s, 0);
        h1_sweep(L, R, g, f1_l, y_j, s, 1);
        h2_sweep(L, R, g, c_l, z_j, s, 0, 0);
        h2_sweep(L, R, g, c_l, z_j, s, 0, 1);
        h2_sweep(L, R, g, c_l, z_j, s, 1, 0);
        h2_sweep(L, R, g, c_l, z_j, s, 1, 1);
        cn_all(L, R, g, h_diag, h_off, tau);
        h2_sweep(L, R, g, c_l, z_j, s, 1, 1);
        h2_sweep(L, R, g, c_l, z_j, s, 1, 0);
        h2_sweep(L, R, g, c_l, z_j, s, 0, 1);
        h2_sweep(L, R, g, c_l, z_j, s, 0, 0);
        h1_sweep(L, R, g, f1_l, y_j, s, 1);
        h1_sweep(L, R, g, f1_l, y_j, s, 0);
        apply_mask(L, R, g, mask);
    }
}

/* ---- a9: LineMesh CN, length gauge; batch of independent sims --------------- */
/* evolution_methods.py:49-77 with H = H0 + diag(-q z E(t_{n+1})) (mesh_operators.py:271-298, :320-327).
   g: [batch][Z]; fields: [nsteps][batch]. */
void ora_line_cn_len_steps(long batch, long Z, c128 *g, const c128 *h_diag, const double *h_off, const double *w_z,
                           const double *mask, long nsteps, const double *taus, const double *fields)
{
#pragma omp parallel
    {
        c128 *work = (c128 *)malloc((size_t)(7 * Z) * sizeof(c128));
        c128 *hd = work + 6 * Z;
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < batch; ++b) {
            c128 *gb = g + b * Z;
            for (long n = 0; n < nsteps; ++n) {
                double e = fields[n * batch + b];
                for (long k = 0; k < Z; ++k) hd[k] = h_diag[k] + e * w_z[k];
                cn_row(Z, gb, hd, h_off, taus[n], work);
                for (long k = 0; k < Z; ++k) gb[k] *= mask[k];
            }
        }
        free(work);
    }
}

/* SO on LineMesh: length gauge phase (mesh_operators.py:329-341) */
void ora_line_so_len_steps(long batch, long Z, c128 *g, const c128 *h_diag, const double *h_off, const double *w_z,
                           const double *mask, long nsteps, const double *taus, const double *fields)
{
#pragma omp parallel
    {
        c128 *work = (c128 *)malloc((size_t)(6 * Z) * sizeof(c128));
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < batch; ++b) {
            c128 *gb = g + b * Z;
            for (long n = 0; n < nsteps; ++n) {
                double s = taus[n] * fields[n * batch + b];
                for (long k = 0; k < Z; ++k) gb[k] *= cexp(-I * s * w_z[k]);
                cn_row(Z, gb, h_diag, h_off, taus[n], work);
                for (long k = 0; k < Z; ++k) gb[k] *= cexp(-I * s * w_z[k]) * mask[k];
            }
        }
        free(work);
    }
}

static void line_vel_sweep(long Z, c128 *g, double th, int parity)
{
    double c = cos(th), sn = sin(th);
    for (long k = parity; k < Z - 1; k += 2) {                /* mesh_operators.py:384-427 */
        c128 lo = g[k], hi = g[k + 1];
        g[k] = c * lo + sn * hi;
        g[k + 1] = -sn * lo + c * hi;
    }
}

void ora_line_so_vel_steps(long batch, long Z, c128 *g, const c128 *h_diag, const double *h_off, double v_pref,
                           const double *mask, long nsteps, const double *taus, const double *fields)
{
#pragma omp parallel
    {
        c128 *work = (c128 *)malloc((size_t)(6 * Z) * sizeof(c128));
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < batch; ++b) {
            c128 *gb = g + b * Z;
            for (long n = 0; n < nsteps; ++n) {
                double th = taus[n] * fields[n * batch + b] * v_pref;
                line_vel_sweep(Z, gb, th, 0);
                line_vel_sweep(Z, gb, th, 1);
                cn_row(Z, gb, h_diag, h_off, taus[n], work);
                line_vel_sweep(Z, gb, th, 1);
                line_vel_sweep(Z, gb, th, 0);
                for (long k = 0; k < Z; ++k) gb[k] *= mask[k];
            }
        }
        free(work);
    }
}

/* ---- a11: observables ------------------------------------------------------- */
double ora_norm(long n, const c128 *g, double ipm)
{
    double acc = 0.0;
    for (long i = 0; i < n; ++i) acc += creal(g[i]) * creal(g[i]) + cimag(g[i]) * cimag(g[i]);
    return acc * ipm;
}

int ora_num_threads(void)
{
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}
