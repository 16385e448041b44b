// ionization_b200 -- field set-up on the device: the per-step scalars of a whole scan in two kernels.
//
// The reference evaluates, for every time step of every simulation, E(t) of the pulse (potentials/pulses.py:929-940: Sinc
// envelope x carrier x time window, windows.py:129-168) and, in the velocity gauge, A(t_n) = -simps(E(times[:n+1]), times[:n+1])
// from scratch (mesh_operators.py:1184-1186, pulses.py:58-77): O(n) per step, O(n^2) per run, per member.  Here
//   k_sinc_field    one thread per (time sample, pulse):  E_b(t_n + offset)
//   k_prefix_simps  one thread per pulse: the old-scipy Simpson rule (even='avg') for EVERY prefix of the samples in one O(n)
//                   sweep -- two running Simpson sums (over [0..m] for even m, over [1..m] for odd m) combined with the two end
//                   trapezoids exactly as scipy <= 1.10 combines them (the host twin is coefficients.prefix_simps).
// Arrays are [n][n_pulses] (pulse fastest): coalesced in both kernels, and the layout ion_sim_step takes for `fields`.
#pragma once
#include "common.cuh"

namespace ion {

struct SincPulseParams {  // one row of the host's [n_pulses][8] parameter table
    double amplitude, delta_omega, omega_carrier, phase, pulse_center, window_time, window_width, window_center;
};

// LogisticWindow.__call__ (windows.py:129-168); window_width <= 0: no window
ION_DEVINL double logistic_window(double t, double wt, double ww, double wc)
{
    if (!(ww > 0.0)) return 1.0;
    const double tau = t - wc;
    return 1.0 / (1.0 + exp(-(tau + wt) / ww)) - 1.0 / (1.0 + exp(-(tau - wt) / ww));
}

// out[n][b] = E_b(times[n] + offset), n = 0 .. n_times - 1
__global__ void k_sinc_field(const double *__restrict__ times, double offset, int n_times, const SincPulseParams *__restrict__ pulses, int n_pulses,
                             double *__restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (b >= n_pulses || n >= n_times) return;
    const SincPulseParams p = pulses[b];
    const double t = times[n] + offset;
    const double tau = t - p.pulse_center;
    const double x = p.delta_omega * tau / 2.0;
    const double env = (x == 0.0) ? 1.0 : sin(x) / x;  // sinc (pulses.py:594-596)
    out[(size_t)n * n_pulses + b] = env * cos(p.omega_carrier * tau + p.phase) * p.amplitude * logistic_window(t, p.window_time, p.window_width, p.window_center);
}

ION_DEVINL double simpson_panel(double y0, double y1, double y2, double h0, double h1)
{
    const double hsum = h0 + h1;
    return hsum / 6.0 * (y0 * (2.0 - h1 / h0) + y1 * hsum * hsum / (h0 * h1) + y2 * (2.0 - h0 / h1));
}

// out[n - 1][b] = sign * simps(y[0..n][b], x[0..n]) for n = 1 .. n_times - 1   (sign = -1: the vector potential)
__global__ void k_prefix_simps(const double *__restrict__ y, const double *__restrict__ x, int n_times, int n_pulses, double sign, double *__restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_pulses || n_times < 2) return;
    double s_even = 0.0, s_odd = 0.0;
    double y2 = y[b], y1 = 0.0, y0 = 0.0;  // y2: newest sample
    const double h_first = x[1] - x[0];
    double first_trap = 0.0;
    for (int n = 1; n < n_times; ++n) {
        y0 = y1;
        y1 = y2;
        y2 = y[(size_t)n * n_pulses + b];
        double val;
        if (n == 1) {
            first_trap = 0.5 * h_first * (y1 + y2);
            val = first_trap;
        } else {
            const double h0 = x[n - 1] - x[n - 2], h1 = x[n] - x[n - 1];
            if ((n & 1) == 0) {
                s_even += simpson_panel(y0, y1, y2, h0, h1);
                val = s_even;
            } else {
                s_odd += simpson_panel(y0, y1, y2, h0, h1);
                const double last_trap = 0.5 * h1 * (y2 + y1);
                val = 0.5 * ((last_trap + s_even) + (first_trap + s_odd));
            }
        }
        out[(size_t)(n - 1) * n_pulses + b] = sign * val;
    }
}

}  // namespace ion
