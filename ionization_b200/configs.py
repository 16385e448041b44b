"""The synthetic workloads of BASELINE.json ``configs`` / SURVEY.md section 8(d), as dictionaries of hot-path inputs
(same keys as tests/golden/*.npz) built with this package's own host-side builders.

Common inputs: dt = 1 as; hydrogen 1s initial state; Coulomb potential; analytic test states n <= 3;
RadialCosineMask(0.8 r_bound, r_bound, smoothness 8); SincPulse(pulse_width 200 as, default omega_min) in a
LogisticWindow(4 pw, 0.2 pw); t in [-5 pw, 5 pw] -> 2000 steps.
"""
import numpy as np

from . import coefficients as C
from . import potentials as P
from . import states as S
from . import units as u


def hydrogen_test_states(n_max=3, l_bound=None):
    out = [S.HydrogenBoundState(n, l) for n in range(1, n_max + 1) for l in range(n)]
    if l_bound is not None:
        out = [s for s in out if s.l < l_bound]
    return out


def sinc_pulse(pulse_width=200 * u.asec, fluence=1 * u.Jcm2, phase=0.0):
    return P.SincPulse(
        pulse_width=pulse_width, fluence=fluence, phase=phase, window=P.LogisticWindow(window_time=4 * pulse_width, window_width=0.2 * pulse_width)
    )


def radial_g(state, r, delta_r):
    """SphericalHarmonicMesh.get_radial_g_for_state (mesh/meshes.py:1090-1097): g = r R(r), normalised on the mesh."""
    g = np.asarray(state.radial_function(r) * r, dtype=np.complex128)
    g = g / np.sqrt(np.real(np.sum(np.conj(g) * g)) * delta_r)
    return g * state.amplitude


def spherical_harmonic_problem(*, r_bound, r_points, l_bound, gauge="LEN", pulse=None, pulse_width=200 * u.asec, time_step=1 * u.asec,
                               time_initial=None, time_final=None, n_steps=None, initial_state=None, test_states=None, mask=True):
    r, delta_r = C.sh_r_grid(r_bound, r_points)
    q, m = u.electron_charge, u.electron_mass_reduced
    pulse = pulse if pulse is not None else sinc_pulse(pulse_width)
    t0 = -5 * pulse_width if time_initial is None else time_initial
    t1 = 5 * pulse_width if time_final is None else time_final
    times = C.time_grid(t0, t1, time_step)
    if n_steps is not None:
        times = times[: n_steps + 1]
    kind = "sh_len_so" if gauge == "LEN" else "sh_vel_so"
    V = P.CoulombPotential(charge=u.proton_charge)(r=r, test_charge=q)
    h_diag, h_off = C.sh_hamiltonian(r, delta_r, l_bound, V)
    initial_state = initial_state or S.HydrogenBoundState(1, 0)
    test_states = sorted(test_states if test_states is not None else hydrogen_test_states(3, l_bound))
    g0 = np.zeros((l_bound, r_points), dtype=np.complex128)
    g0[initial_state.l] = radial_g(initial_state, r, delta_r)
    prob = dict(
        kind=kind, L=l_bound, R=r_points, r=r, delta_r=delta_r, h_diag=h_diag, h_off=h_off, g0=g0, times=times,
        taus=C.taus_from_times(times), fields=C.field_series(kind, pulse, times, time_step), time_step=time_step,
        mask=(P.RadialCosineMask(0.8 * r_bound, r_bound, 8)(r=r).astype(np.float64) if mask else np.ones(r_points)),
        state_l=np.array([s.l for s in test_states], dtype=np.int64),
        state_rows=np.array([radial_g(s, r, delta_r) for s in test_states]),
        state_bound=np.array([s.bound for s in test_states]),
        initial_state_index=test_states.index(initial_state) if initial_state in test_states else -1,
        test_charge=q, test_mass=m,
    )
    if gauge == "LEN":
        prob["c_l"], prob["x_j"] = C.sh_len_coupling(r, l_bound, q)
    else:
        prob["c_l"], prob["f1_l"], prob["y_j"], prob["z_j"] = C.sh_vel_coupling(r, delta_r, l_bound, q, m)
    return prob


def config1(gauge="LEN", **kw):
    """configs[0]: hydrogen 1s, r_bound=100 a0, r_points=500, l_bound=50, Sinc pulse, length-gauge split operator"""
    return spherical_harmonic_problem(r_bound=100 * u.bohr_radius, r_points=500, l_bound=50, gauge=gauge, **kw)


def config3(gauge="VEL", **kw):
    """configs[2]: hydrogen 1s, r_bound=250 a0, r_points=2000, l_bound=500, velocity gauge, single sim"""
    return spherical_harmonic_problem(r_bound=250 * u.bohr_radius, r_points=2000, l_bound=500, gauge=gauge, **kw)


def config4_member(gauge="LEN", **kw):
    """configs[3], one member: r_bound=100 a0, r_points=1000, l_bound=200, length gauge"""
    return spherical_harmonic_problem(r_bound=100 * u.bohr_radius, r_points=1000, l_bound=200, gauge=gauge, **kw)


def config4_fields(problem, n_fluence=64, n_phase=64, pulse_width=200 * u.asec):
    """fluence x CEP scan of configs[3]: geomspace(0.01, 20) J/cm^2 x linspace(0, 2 pi) -> fields [n_steps, batch]"""
    flu = np.geomspace(0.01, 20, n_fluence) * u.Jcm2
    ph = np.linspace(0, u.twopi, n_phase, endpoint=False)
    cols = []
    for f in flu:
        for p in ph:
            cols.append(C.field_series(str(problem["kind"]), sinc_pulse(pulse_width, f, p), problem["times"], problem["time_step"]))
    return np.ascontiguousarray(np.array(cols).T)


def line_problem(*, z_bound, z_points, kind="line_len_cn", pulse=None, pulse_width=200 * u.asec, time_step=1 * u.asec, n_steps=1000, mask=True):
    """LineMesh inputs (configs[1]): Gaussian well -10 eV / 5 a0 (dev/potentials/gaussian_well.py:19-21 of the reference),
    electron mass, variational ground state, t in [-n_steps/2, n_steps/2] dt."""
    z, dz = C.line_z_grid(z_bound, z_points)
    q, m = u.electron_charge, u.electron_mass
    well = P.GaussianPotential(potential_extrema=-10 * u.eV, width=5 * u.bohr_radius)
    state = S.GaussianWellState.from_potential(well, m)
    pulse = pulse if pulse is not None else sinc_pulse(pulse_width)
    times = C.time_grid(-n_steps / 2 * time_step, n_steps / 2 * time_step, time_step)[: n_steps + 1]
    h_diag, h_off = C.line_hamiltonian(z, dz, well(r=z), m)
    w_z, v_pref = C.line_coupling(z, dz, q, m)
    g0 = np.asarray(state(z), dtype=np.complex128)
    g0 = g0 / np.sqrt(np.real(np.sum(np.conj(g0) * g0)) * dz)
    return dict(
        kind=kind, Z=z_points, z=z, delta_z=dz, h_diag=h_diag, h_off=h_off, w_z=w_z, v_pref=v_pref, g0=g0, times=times, taus=C.taus_from_times(times),
        fields=C.field_series(kind, pulse, times, time_step), time_step=time_step,
        mask=(P.RadialCosineMask(0.8 * z_bound, z_bound, 8)(r=z).astype(np.float64) if mask else np.ones(z_points)),
        state_rows=g0[None, :].copy(), initial_state_index=0, test_charge=q, test_mass=m,
    )


def config2(**kw):
    """configs[1]: LineMesh 1D Gaussian well, 2^16 points, Crank-Nicolson length gauge (batch of Sinc pulses: scan_fields)"""
    return line_problem(z_bound=2000 * u.bohr_radius, z_points=2 ** 16, kind="line_len_cn", **kw)


def scan_fields(problem, fluences_jcm2, phases, pulse_width=200 * u.asec):
    """fluence x CEP scan (ionization_scans/scan_mesh.py:40-68): fields [n_steps, len(fluences) * len(phases)]"""
    cols = []
    for f in fluences_jcm2:
        for p in phases:
            cols.append(C.field_series(str(problem["kind"]), sinc_pulse(pulse_width, f * u.Jcm2, p), problem["times"], problem["time_step"]))
    return np.ascontiguousarray(np.array(cols).T)
