"""Host-side builders of the hot path's inputs: the coefficient vectors and per-step scalars the CUDA engine
consumes.  This is the part of the reference's ``MeshOperators`` classes that survives once the operators are
no longer materialised as sparse matrices every step (mesh/mesh_operators.py).

Everything here is O(R + L) or O(L R) numpy work done once per simulation (or once per time step for a scalar);
nothing here touches the wavefunction.  Checked against the reference's own matrices in
tests/test_host_inputs.py (fixtures dumped from the live reference objects).
"""
import numpy as np

from . import units as u


# ---- SphericalHarmonicMesh -------------------------------------------------------------------
def sh_alpha(j):
    """mesh_operators.py:841-843"""
    j = np.asarray(j, dtype=np.float64)
    x = j ** 2 + 2 * j
    return (x + 1) / (x + 0.75)


def sh_beta(j):
    """mesh_operators.py:845-847"""
    j = np.asarray(j, dtype=np.float64)
    x = 2 * j ** 2 + 2 * j
    return (x + 1) / (x + 0.5)


def sh_c_l(l):
    """mesh_operators.py:853-855"""
    l = np.asarray(l, dtype=np.float64)
    return (l + 1) / np.sqrt((2 * l + 1) * (2 * l + 3))


def sh_r_grid(r_bound, r_points):
    """mesh/meshes.py:1009-1011: r_j = (j + 1/2) delta_r with delta_r = r_bound / (r_points - 1)."""
    r = np.linspace(0, r_bound, r_points)
    delta_r = r[1] - r[0]
    return r + delta_r / 2, delta_r


def sh_hamiltonian(r, delta_r, l_bound, potential_r, hydrogen_zero_angular_momentum_correction=True, l_begin=0):
    """(h_diag [L, R] complex128, h_off [R-1] float64): field-free radial Hamiltonian per channel.

    kinetic_energy_from_lagrangian (mesh_operators.py:889-928; always electron_mass_reduced, :892-894) plus the
    potential on the diagonal (internal_hamiltonian, :244-269).  ``l_begin``: first channel (l-block shards).
    """
    R = len(r)
    j = np.arange(R)
    pre = -(u.hbar ** 2) / (2 * u.electron_mass_reduced * delta_r ** 2)
    beta = np.tile(sh_beta(j), (l_bound, 1)).astype(np.complex128)
    if hydrogen_zero_angular_momentum_correction and l_begin == 0:
        dr = delta_r / u.bohr_radius
        beta[0, 0] += dr * (1 + dr) / 8  # :901-903
    h_diag = beta * (-2 * pre)
    l = np.arange(l_begin, l_begin + l_bound)
    h_diag = h_diag + ((u.hbar ** 2) / (2 * u.electron_mass_reduced)) * (l * (l + 1))[:, None] / (r[None, :] ** 2)
    h_diag = h_diag + np.asarray(potential_r)[None, :]
    h_off = (sh_alpha(j[:-1]) * pre).astype(np.float64)
    return h_diag, h_off


def sh_len_coupling(r, l_total, test_charge):
    """angle(l, j) = tau * E * c_l * x_j,  x_j = -q r_j  (mesh_operators.py:988-1006, :1043)"""
    return sh_c_l(np.arange(l_total - 1)), -test_charge * np.asarray(r)


def sh_vel_coupling(r, delta_r, l_total, test_charge, test_mass):
    """theta1 = tau*A * f1_l * y_j, theta2 = tau*A * c_l * z_j  (mesh_operators.py:1143-1178; test_mass, not the
    reduced mass: SURVEY App. B-3)"""
    l = np.arange(l_total - 1)
    c_l = sh_c_l(l)
    f1_l = c_l * (l + 1)
    y_j = u.hbar * (test_charge / test_mass) / np.asarray(r)
    z_j = u.hbar * (test_charge / test_mass) / (2 * delta_r) * sh_alpha(np.arange(len(r) - 1))
    return c_l, f1_l, y_j, z_j


# ---- LineMesh --------------------------------------------------------------------------------
def line_z_grid(z_bound, z_points):
    """mesh/meshes.py:285-288"""
    z = np.linspace(-z_bound, z_bound, z_points)
    return z, np.abs(z[1] - z[0])


def line_hamiltonian(z, delta_z, potential_z, test_mass):
    """mesh_operators.py:310-318 + :244-269"""
    pre = -(u.hbar ** 2) / (2 * test_mass * delta_z ** 2)
    h_diag = (-2 * pre) * np.ones(len(z), dtype=np.complex128) + np.asarray(potential_z)
    h_off = pre * np.ones(len(z) - 1, dtype=np.float64)
    return h_diag, h_off


def line_coupling(z, delta_z, test_charge, test_mass):
    """w_z = -q z (length gauge, :320-327), v_pref = hbar (q/m) / (2 dz) (velocity gauge, :358-375)"""
    return -test_charge * np.asarray(z), u.hbar * (test_charge / test_mass) / (2 * delta_z)


# ---- time grid and per-step scalars ------------------------------------------------------------
def time_grid(time_initial, time_final, time_step, spec=None):
    """MeshSimulation.get_times (mesh/sims.py:198-220)"""
    if not callable(time_step):
        total_time = time_final - time_initial
        return np.linspace(time_initial, time_final, int(total_time / time_step) + 1)
    t = time_initial
    times = [t]
    while t < time_final:
        t += time_step(t, spec)
        if t > time_final:
            t = time_final
        times.append(t)
    return np.array(times)


def taus_from_times(times):
    """tau_n = (t_n - t_{n-1}) / (2 hbar)  (mesh/sims.py:319-321, evolution_methods.py:92)"""
    return np.diff(np.asarray(times, dtype=np.float64)) / (2 * u.hbar)


def _simpson_panel(y0, y1, y2, h0, h1):
    """one non-uniform Simpson panel over (x0, x1, x2); the summand of old scipy's _basic_simps"""
    hsum = h0 + h1
    return hsum / 6.0 * (y0 * (2 - h1 / h0) + y1 * hsum * hsum / (h0 * h1) + y2 * (2 - h0 / h1))


def prefix_simps(y, x):
    """out[n] = potentials.simps(y[:n+1], x[:n+1]) for every n >= 1, in O(N) total.

    Two running Simpson sums are kept: over [0..m] for even m and over [1..m] for odd m; the even-sample-count
    prefixes combine them with the two end trapezoids exactly as old scipy's ``even='avg'`` does.
    """
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    N = len(y)
    out = np.zeros(N)
    h = np.diff(x)
    s_even = 0.0  # Simpson over [0..m], m even
    s_odd = 0.0  # Simpson over [1..m], m odd
    first_trap = 0.5 * h[0] * (y[0] + y[1]) if N > 1 else 0.0
    for n in range(1, N):
        if n % 2 == 0:
            s_even += _simpson_panel(y[n - 2], y[n - 1], y[n], h[n - 2], h[n - 1])
            out[n] = s_even
        else:
            if n >= 3:
                s_odd += _simpson_panel(y[n - 2], y[n - 1], y[n], h[n - 2], h[n - 1])
            if n == 1:
                out[n] = first_trap
            else:
                last_trap = 0.5 * h[n - 1] * (y[n] + y[n - 1])
                out[n] = 0.5 * ((last_trap + s_even) + (first_trap + s_odd))
    return out


def vector_potential_series(electric_potential, times):
    """A_n = electric_potential.get_vector_potential_amplitude_numeric(times[:n+1]) for n = 1..N-1.

    The reference recomputes this from scratch every step -- O(n) each, O(n^2) per run (mesh_operators.py:1184-1186,
    SURVEY 3.3).  For this package's own pulse classes the field is sampled once and the old-scipy Simpson rule is
    evaluated for all prefixes incrementally (O(n) total); a foreign pulse object (e.g. the reference's own) is
    simply asked step by step, as the reference does.
    """
    from . import potentials

    times = np.asarray(times, dtype=np.float64)
    N = len(times)
    native = isinstance(electric_potential, (potentials.UniformLinearlyPolarizedElectricPotential, potentials.PotentialEnergySum))
    if not native:
        return np.array([electric_potential.get_vector_potential_amplitude_numeric(times[: n + 1]) for n in range(1, N)], dtype=np.float64)
    y = np.asarray(electric_potential.get_electric_field_amplitude(times), dtype=np.float64) * np.ones(N)
    return -prefix_simps(y, times)[1:]


def field_series(program: str, electric_potential, times, time_step):
    """The scalar the reference samples for each step n -> n+1:
    SH length gauge E(t_{n+1} + time_step/2) (mesh_operators.py:1011-1013), Line length gauge E(t_{n+1}) (:321-323),
    velocity gauge A over times[0..n+1] (:1184-1186, :373-375)."""
    times = np.asarray(times, dtype=np.float64)
    if program in ("sh_len_so", "sh_len_adi"):
        ts = float(time_step) if not callable(time_step) else None
        if ts is None:
            raise ValueError("SphericalHarmonic length gauge samples E at t + spec.time_step/2: time_step must be a number")
        return np.asarray(electric_potential.get_electric_field_amplitude(times[1:] + ts / 2), dtype=np.float64) * np.ones(len(times) - 1)
    if program in ("line_len_cn", "line_len_so"):
        return np.asarray(electric_potential.get_electric_field_amplitude(times[1:]), dtype=np.float64) * np.ones(len(times) - 1)
    if program in ("sh_vel_so", "line_vel_so"):
        return vector_potential_series(electric_potential, times)
    raise ValueError(program)


def sinc_pulse_table(pulses):
    """[n, 8] parameter table of plain windowed Sinc pulses for the device field set-up (include/ionization_b200.h:
    ion_sinc_pulse_fields), or None if any pulse is something else (a sum, a DC-corrected pulse, another window ...)"""
    from . import potentials

    rows = []
    for p in pulses:
        if type(p) is not potentials.SincPulse:
            return None
        w = p.window
        if type(w) is potentials.LogisticWindow:
            win = (w.window_time, w.window_width, w.window_center)
        elif type(w) is potentials.NoTimeWindow:
            win = (0.0, 0.0, 0.0)
        else:
            return None
        rows.append((p.amplitude, p.delta_omega, p.omega_carrier, p.phase, p.pulse_center) + win)
    return np.ascontiguousarray(rows, dtype=np.float64)


def field_series_batch(program: str, pulses, times, time_step, device=None):
    """``field_series`` for every pulse of a scan -> [n_steps, n_pulses].  Plain windowed Sinc pulses are evaluated on the GPU
    (two kernels for the whole scan, csrc/fields.cuh) when ``device`` is given; anything else, pulse by pulse on the host."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    table = sinc_pulse_table(pulses) if device is not None else None
    if table is None or len(times) < 2:
        return np.ascontiguousarray(np.array([field_series(program, p, times, time_step) for p in pulses]).T)
    import ctypes

    from . import _native as nat

    if program in ("sh_len_so", "sh_len_adi"):
        kind, offset = 0, float(time_step) / 2
    elif program in ("line_len_cn", "line_len_so"):
        kind, offset = 0, 0.0
    elif program in ("sh_vel_so", "line_vel_so"):
        kind, offset = 1, 0.0
    else:
        raise ValueError(program)
    out = np.empty((len(times) - 1, len(table)), dtype=np.float64)
    nat.check(nat.load().ion_sinc_pulse_fields(int(device), kind, len(times), nat.ptr(times), ctypes.c_double(offset), len(table), nat.ptr(table), nat.ptr(out)),
              "ion_sinc_pulse_fields")
    return out
