"""ORACLE TEST INFRASTRUCTURE -- container-only fixture generator.

Runs the UNMODIFIED reference (``/root/reference`` under ``oracle/shim``) and
dumps, for each case, (1) the inputs of the hot path exactly as the reference
built them (Hamiltonian diagonals out of its own ``dia_matrix`` objects, the
per-step field scalars out of its own pulse objects, mask, initial ``g``,
test-state rows) and (2) its outputs (final ``g``, norm and inner-product
series).  The ``.npz`` files are committed under ``tests/golden/``; the oracle
(``oracle/restate.py``), the host-side coefficient builders and the CUDA engine
are all checked against them.

    python -m oracle.make_golden [--only NAME] [--big]

``--big`` also regenerates the six known-answer runs of
dev/meshes/mesh_refactoring_helper.py:30-86,204-251 (minutes of CPU).
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import restate  # noqa: E402
from oracle.refenv import import_reference  # noqa: E402

ion = import_reference()
import simulacra.units as u  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CONSTANTS = dict(
    hbar=u.hbar,
    electron_mass=u.electron_mass,
    electron_mass_reduced=u.electron_mass_reduced,
    electron_charge=u.electron_charge,
    proton_charge=u.proton_charge,
    bohr_radius=u.bohr_radius,
    coulomb_constant=u.coulomb_constant,
    epsilon_0=u.epsilon_0,
    c=u.c,
    asec=u.asec,
    eV=u.eV,
    Jcm2=u.Jcm2,
    atomic_electric_field=u.atomic_electric_field,
    rydberg=u.rydberg,
)


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (d if d > 0 else 1.0))


# --------------------------------------------------------------------------
# input extraction from live reference objects
# --------------------------------------------------------------------------
def _sh_inputs(sim):
    spec, mesh = sim.spec, sim.mesh
    L, R = mesh.mesh_shape
    ops = spec.operators
    (h0,) = ops.internal_hamiltonian(mesh).operators  # mesh_operators.py:244-269
    dia = h0.matrix
    offsets = list(dia.offsets)
    diag = dia.data[offsets.index(0)].reshape(L, R)
    sup = dia.data[offsets.index(1)][1:]  # scipy dia: data[k][j] sits in column j
    sub = dia.data[offsets.index(-1)][:-1]
    assert np.array_equal(sup, sub)
    off = sup.reshape(-1)
    off_blocks = np.concatenate([off, [0]]).reshape(L, R)
    assert np.all(off_blocks[:, -1] == 0), "off-diagonal must vanish at l-block edges"
    assert np.all(off_blocks[:, :-1] == off_blocks[0, :-1]), "off-diagonal must be l-independent"
    h_off = np.real(off_blocks[0, :-1]).copy()
    assert np.all(np.imag(off_blocks[0, :-1]) == 0)

    l = np.arange(L)
    c_l = restate.sh_c_l(l[:-1])
    q = spec.test_charge
    x_j = -q * mesh.r
    out = dict(
        R=R,
        L=L,
        r=mesh.r,
        delta_r=mesh.delta_r,
        h_diag=diag.copy(),
        h_off=h_off,
        c_l=c_l,
        x_j=x_j,
        mask=np.broadcast_to(np.asarray(spec.mask(r=mesh.r), dtype=np.float64), (R,)).copy(),
        g0=mesh.g.copy(),
        times=sim.times.copy(),
        taus=np.diff(sim.times) / (2 * u.hbar),
        time_step=float(spec.time_step),
        test_charge=q,
        test_mass=spec.test_mass,
    )
    # cross-check the coupling vectors against the reference's own matrices
    if isinstance(ops, ion.mesh.SphericalHarmonicVelocityGaugeOperators):
        h1, h2 = ops.interaction_hamiltonian_matrices_without_field(mesh)
        f1_l = c_l * (l[:-1] + 1)
        y_j = u.hbar * (q / spec.test_mass) / mesh.r
        z_j = u.hbar * (q / spec.test_mass) / (2 * mesh.delta_r) * restate.sh_alpha(np.arange(R - 1))
        # h1 is L-wrapped: flat k = j*L + l, superdiagonal data[-1][1:]
        a1 = (h1.data[-1][1:] * (-1j)).real  # length N-1
        full = np.concatenate([a1, [0]]).reshape(R, L).T  # (L, R)
        assert _rel(f1_l[:, None] * y_j[None, :], full[:-1]) < 1e-14
        assert np.all(full[-1] == 0)
        a2 = (h2.data[-1][R + 1 :] * (-1j)).real  # mesh_operators.py:1252
        full2 = np.concatenate([a2, np.zeros(R + 1)]).reshape(L, R)
        assert _rel(c_l[:, None] * z_j[None, :], full2[:-1, :-1]) < 1e-14
        out.update(f1_l=f1_l, y_j=y_j, z_j=z_j)
    else:
        lint = ops.interaction_hamiltonian_matrices_without_field(mesh)
        a = np.real(lint.data[0][:-1])
        full = np.concatenate([a, [0]]).reshape(R, L).T
        assert _rel(c_l[:, None] * x_j[None, :], full[:-1]) < 1e-14
        assert np.all(full[-1] == 0)
    return out


def _sh_fields(sim, kind):
    """per-step scalar exactly as the reference samples it (sim.time is already t_{n+1}
    when evolve() runs, mesh/sims.py:317-319)."""
    spec, times = sim.spec, sim.times
    pot = spec.electric_potential
    if kind in ("sh_len_so", "sh_len_adi"):
        return np.array([pot.get_electric_field_amplitude(times[n] + spec.time_step / 2) for n in range(1, len(times))], dtype=np.float64)
    return np.array([pot.get_vector_potential_amplitude_numeric(times[: n + 1]) for n in range(1, len(times))], dtype=np.float64)


def _states_sh(sim):
    spec, mesh = sim.spec, sim.mesh
    ls, rows, names, bound = [], [], [], []
    for s in spec.test_states:
        ls.append(int(s.l))
        rows.append(mesh.get_radial_g_for_state(s))
        names.append(str(s))
        bound.append(bool(s.bound))
    init = spec.test_states.index(spec.initial_state) if spec.initial_state in spec.test_states else -1
    return dict(state_l=np.array(ls), state_rows=np.array(rows), state_names=np.array(names), state_bound=np.array(bound), initial_state_index=init)


def _outputs(sim):
    out = dict(g_final=sim.mesh.g.copy(), norm=sim.data.norm.copy(), data_times=sim.data.times.copy())
    ips = np.array([sim.data.inner_products[s] for s in sim.spec.test_states]).T  # (n_data, n_states)
    out["inner_products"] = ips
    return out


def dump_sh(name, kind, spec_kwargs, *, store=1, keep_g=True, extra_datastores=False):
    ops, method = {
        "sh_len_so": (ion.mesh.SphericalHarmonicLengthGaugeOperators, ion.mesh.SplitInteractionOperator),
        "sh_vel_so": (ion.mesh.SphericalHarmonicVelocityGaugeOperators, ion.mesh.SplitInteractionOperator),
        "sh_len_adi": (ion.mesh.SphericalHarmonicLengthGaugeOperators, ion.mesh.AlternatingDirectionImplicit),
    }[kind]
    kw = dict(spec_kwargs)
    if extra_datastores:
        D = ion.mesh
        kw["datastores"] = [
            D.Fields(),
            D.Norm(),
            D.InnerProducts(),
            D.InternalEnergyExpectationValue(),
            *([D.TotalEnergyExpectationValue()] if kind != "sh_vel_so" else []),  # VEL: SumOfOperators of raw matrices (mesh_operators.py:1188) cannot be applied
            D.ZExpectationValue(),
            D.RExpectationValue(),
            D.NormWithinRadius(radii=[r * u.bohr_radius for r in (5, 10, 20)]),
        ]
    t0 = time.perf_counter()
    sim = ion.mesh.SphericalHarmonicSpecification(name, operators=ops(), evolution_method=method(), store_data_every=store, **kw).to_sim()
    d = dict(kind=kind, **_sh_inputs(sim), **_states_sh(sim))
    d["fields"] = _sh_fields(sim, kind)
    nbl = []
    # NormBySphericalHarmonic.init reads self.spec before it is set (mesh/data.py:441-446), so the
    # datastore cannot be attached in the reference; sample mesh.norm_by_l() (meshes.py:1133-1136) directly.
    sim.run(callback=(lambda s: nbl.append(s.mesh.norm_by_l())) if extra_datastores else None)
    d.update(_outputs(sim))
    if extra_datastores:
        d["internal_energy"] = sim.data.internal_energy_expectation_value.copy()
        if kind != "sh_vel_so":
            d["total_energy"] = sim.data.total_energy_expectation_value.copy()
        d["z_expectation"] = sim.data.z_expectation_value.copy()
        d["r_expectation"] = sim.data.r_expectation_value.copy()
        d["norm_within_radii"] = np.array(sorted(sim.data.norm_within_radius.keys()))
        d["norm_within_radius"] = np.array([sim.data.norm_within_radius[r] for r in sorted(sim.data.norm_within_radius.keys())]).T
        d["norm_by_l"] = np.array(nbl)[sim.data_mask]
        d["electric_field_amplitude"] = sim.data.electric_field_amplitude.copy()
        d["vector_potential_amplitude"] = sim.data.vector_potential_amplitude.copy()
        # field scalar the reference's total Hamiltonian uses at each data time (mesh_operators.py:1011-1013)
        d["efield_half_at_data_times"] = np.array(
            [sim.spec.electric_potential.get_electric_field_amplitude(t + sim.spec.time_step / 2) for t in sim.data.times]
        )
    d["initial_state_overlap_final"] = float(sim.data.initial_state_overlap[-1])
    if not keep_g:
        del d["g_final"]
    d.update({f"const_{k}": v for k, v in CONSTANTS.items()})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {kind} L={d['L']} R={d['R']} steps={len(d['taus'])} states={len(d['state_l'])} "
          f"norm_f={d['norm'][-1]:.15f} ov_f={d['initial_state_overlap_final']:.12f} "
          f"({time.perf_counter() - t0:.1f}s, {os.path.getsize(path) / 1e6:.2f} MB)", flush=True)
    return d


def _line_inputs(sim):
    spec, mesh = sim.spec, sim.mesh
    ops = spec.operators
    (h0,) = ops.internal_hamiltonian(mesh).operators
    dia = h0.matrix
    offsets = list(dia.offsets)
    diag = dia.data[offsets.index(0)].copy()
    sup = dia.data[offsets.index(1)][1:]
    sub = dia.data[offsets.index(-1)][:-1]
    assert np.array_equal(sup, sub) and np.all(np.imag(sup) == 0)
    Z = len(mesh.z_mesh)
    return dict(
        Z=Z,
        z=mesh.z_mesh.copy(),
        delta_z=mesh.delta_z,
        h_diag=diag,
        h_off=np.real(sup).copy(),
        w_z=-spec.test_charge * mesh.z_mesh,
        v_pref=u.hbar * (spec.test_charge / spec.test_mass) / (2 * mesh.delta_z),
        mask=np.broadcast_to(np.asarray(spec.mask(r=mesh.r_mesh), dtype=np.float64), (Z,)).copy(),
        g0=np.asarray(mesh.g, dtype=np.complex128).copy(),
        times=sim.times.copy(),
        taus=np.diff(sim.times) / (2 * u.hbar),
        time_step=float(spec.time_step),
        test_charge=spec.test_charge,
        test_mass=spec.test_mass,
    )


def dump_line(name, kind, spec_kwargs, *, store=1):
    ops, method = {
        "line_len_cn": (ion.mesh.LineLengthGaugeOperators, ion.mesh.AlternatingDirectionImplicit),
        "line_len_so": (ion.mesh.LineLengthGaugeOperators, ion.mesh.SplitInteractionOperator),
        "line_vel_so": (ion.mesh.LineVelocityGaugeOperators, ion.mesh.SplitInteractionOperator),
    }[kind]
    t0 = time.perf_counter()
    sim = ion.mesh.LineSpecification(name, operators=ops(), evolution_method=method(), store_data_every=store, **spec_kwargs).to_sim()
    d = dict(kind=kind, **_line_inputs(sim))
    spec, times = sim.spec, sim.times
    pot = spec.electric_potential
    if kind == "line_vel_so":
        d["fields"] = np.array([pot.get_vector_potential_amplitude_numeric(times[: n + 1]) for n in range(1, len(times))])
    else:
        d["fields"] = np.array([pot.get_electric_field_amplitude(times[n]) for n in range(1, len(times))], dtype=np.float64)
    d["state_rows"] = np.array([np.asarray(sim.mesh.get_g_for_state(s), dtype=np.complex128) for s in spec.test_states])
    d["state_names"] = np.array([str(s) for s in spec.test_states])
    d["initial_state_index"] = spec.test_states.index(spec.initial_state)
    sim.run()
    d.update(_outputs(sim))
    d["initial_state_overlap_final"] = float(sim.data.initial_state_overlap[-1])
    d.update({f"const_{k}": v for k, v in CONSTANTS.items()})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {kind} Z={d['Z']} steps={len(d['taus'])} norm_f={d['norm'][-1]:.15f} "
          f"ov_f={d['initial_state_overlap_final']:.12f} ({time.perf_counter() - t0:.1f}s, {os.path.getsize(path) / 1e6:.2f} MB)", flush=True)
    return d


def dump_sh_analysis(name, spec_kwargs, snapshot_indices=(20, 60), theta_points=24):
    """SphericalHarmonicSnapshot plane-wave overlaps (snapshots.py:37-80, meshes.py:1138-1190) taken by the
    reference's own run loop, and the radial probability current density of the final state (meshes.py:1358-1370,
    mesh_operators.py:1106-1127, data.py:496-511).  The reference's DirectionalRadialProbabilityCurrent datastore cannot be
    attached (its store() takes no arguments and meshes.py:1359 omits the mesh argument), so the same statements are
    evaluated here on the reference's objects with the mesh passed explicitly."""
    t0 = time.perf_counter()
    snap_kw = dict(plane_wave_overlap__max_wavenumber=20 * u.per_nm, plane_wave_overlap__wavenumber_points=12, plane_wave_overlap__theta_points=9)
    sim = ion.mesh.SphericalHarmonicSpecification(
        name, operators=ion.mesh.SphericalHarmonicLengthGaugeOperators(), evolution_method=ion.mesh.SplitInteractionOperator(), store_data_every=20,
        theta_points=theta_points, **spec_kwargs,
    ).to_sim()
    d = dict(kind="sh_len_so", **_sh_inputs(sim), **_states_sh(sim))
    d["fields"] = _sh_fields(sim, "sh_len_so")
    # The reference's SphericalHarmonicSnapshot cannot take its free-only overlaps: snapshots.py:73 calls
    # get_g_with_states_removed(bound_states) without the required g (meshes.py:163-165) and raises TypeError.  The same
    # statements (snapshots.py:59-80) are evaluated here on the reference's mesh with g passed, at the snapshot indices.
    thetas = np.linspace(0, u.twopi, snap_kw["plane_wave_overlap__theta_points"])
    wavenumbers = np.delete(np.linspace(0, snap_kw["plane_wave_overlap__max_wavenumber"], snap_kw["plane_wave_overlap__wavenumber_points"] + 1), 0)
    taken = {}

    def cb(s):
        if s.time_index in snapshot_indices:
            m = s.mesh
            g_free = m.get_g_with_states_removed(s.bound_states, m.g)
            taken[s.time_index] = (m.norm(), m.inner_product_with_plane_waves(thetas, wavenumbers, g=None)[2],
                                   m.inner_product_with_plane_waves(thetas, wavenumbers, g=g_free)[2])

    sim.run(callback=cb)
    d.update(_outputs(sim))
    d["snapshot_indices"] = np.array(sorted(taken))
    d["snapshot_thetas"], d["snapshot_wavenumbers"] = thetas, wavenumbers
    for idx, (nrm, ip_all, ip_free) in taken.items():
        d[f"snapshot_{idx}_norm"] = nrm
        d[f"snapshot_{idx}_inner_product_with_plane_waves"] = ip_all
        d[f"snapshot_{idx}_inner_product_with_plane_waves__free_only"] = ip_free
    mesh = sim.mesh
    op = sim.spec.operators.r_probability_current__spatial(mesh)
    g_spatial = mesh.space_g_calc
    grad = np.reshape(op.matrix.dot(g_spatial.flatten("F")), g_spatial.shape, "F")
    density = np.imag(np.conj(g_spatial) * grad)
    theta = mesh.theta_calc
    integrand = density * np.sin(theta) * np.abs(theta[1] - theta[0]) * u.twopi
    up = theta <= u.pi / 2
    d["theta_points"] = theta_points
    d["radial_current_density_final"] = density
    d["radial_current_pos_z_final"] = np.sum(integrand[:, up], axis=1) * mesh.r ** 2
    d["radial_current_neg_z_final"] = np.sum(integrand[:, ~up], axis=1) * mesh.r ** 2
    d["snapshot_kwargs_max_wavenumber"] = snap_kw["plane_wave_overlap__max_wavenumber"]
    d.update({f"const_{k}": v for k, v in CONSTANTS.items()})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: analysis fixture, snapshots at {sorted(taken)} ({time.perf_counter() - t0:.1f}s, {os.path.getsize(path) / 1e6:.2f} MB)", flush=True)
    return d


def dump_line_datastores(name, kind, spec_kwargs):
    """LineMesh with the expectation-value datastores (<z> and the energies on a line mesh were untested)"""
    ops, method = {
        "line_len_cn": (ion.mesh.LineLengthGaugeOperators, ion.mesh.AlternatingDirectionImplicit),
        "line_len_so": (ion.mesh.LineLengthGaugeOperators, ion.mesh.SplitInteractionOperator),
    }[kind]
    D = ion.mesh
    sim = ion.mesh.LineSpecification(
        name, operators=ops(), evolution_method=method(), store_data_every=5,
        datastores=[D.Fields(), D.Norm(), D.InnerProducts(), D.InternalEnergyExpectationValue()], **spec_kwargs,  # total energy: total_hamiltonian is a tuple on a line mesh (mesh_operators.py:271-298), the reference cannot evaluate it
    ).to_sim()
    d = dict(kind=kind, **_line_inputs(sim))
    d["fields"] = np.array([sim.spec.electric_potential.get_electric_field_amplitude(sim.times[n]) for n in range(1, len(sim.times))], dtype=np.float64)
    d["state_rows"] = np.array([np.asarray(sim.mesh.get_g_for_state(s), dtype=np.complex128) for s in sim.spec.test_states])
    d["initial_state_index"] = sim.spec.test_states.index(sim.spec.initial_state)
    # the initial g of a real state is float64, and SumOfOperators.apply accumulates complex terms into zeros_like(g)
    # (mesh_operators.py:127-131): same values, complex dtype, so that the reference can evaluate its expectation values at t_0
    sim.mesh.g = sim.mesh.g.astype(np.complex128)
    # LineLengthGaugeOperators.z returns a bare tuple (mesh_operators.py:347-349), which QuantumMesh.expectation_value cannot
    # apply (meshes.py:212), so ZExpectationValue cannot be attached on a line mesh; the reference's own operator is wrapped in
    # its SumOfOperators and evaluated with its own expectation_value at the data times
    from ionization.mesh import mesh_operators as _mo

    zs = []
    sim.run(callback=lambda s_: zs.append(s_.mesh.expectation_value(None, _mo.SumOfOperators(*s_.spec.operators.z(s_.mesh)))))
    d.update(_outputs(sim))
    d["internal_energy"] = sim.data.internal_energy_expectation_value.copy()
    d["z_expectation"] = np.array(zs)[sim.data_mask]
    d.update({f"const_{k}": v for k, v in CONSTANTS.items()})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {kind} line datastores ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)
    return d


# --------------------------------------------------------------------------
# the cases
# --------------------------------------------------------------------------
def hydrogen_states(nmax=3):
    return [ion.states.HydrogenBoundState(n, l) for n in range(1, nmax + 1) for l in range(n)]


def sinc(pw_as, fluence_jcm2=1.0, phase=0.0, window=True):
    pw = pw_as * u.asec
    kw = {}
    if window:
        kw["window"] = ion.potentials.LogisticWindow(window_time=4 * pw, window_width=0.2 * pw)
    return ion.potentials.SincPulse(pulse_width=pw, fluence=fluence_jcm2 * u.Jcm2, phase=phase, **kw)


def small_sh_kwargs(R, L, steps, pw_as, r_bound=30, fluence=1.0, phase=0.0):
    rb = r_bound * u.bohr_radius
    return dict(
        r_bound=rb,
        r_points=R,
        l_bound=L,
        time_initial=-steps / 2 * u.asec,
        time_final=steps / 2 * u.asec,
        time_step=1 * u.asec,
        electric_potential=sinc(pw_as, fluence, phase, window=False),
        use_numeric_eigenstates=False,
        test_states=hydrogen_states(3) if L >= 3 else hydrogen_states(min(L, 2)),
        mask=ion.potentials.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb, smoothness=8),
    )


def cases(big):
    """yield (name, thunk)"""
    # --- small SH cases, all four (R, L) parities (Appendix B-2/B-6 quirks) ---
    for R, L in ((100, 10), (101, 11), (100, 11), (101, 10)):
        n = f"sh_len_so_{R}x{L}"
        yield n, lambda n=n, R=R, L=L: dump_sh(n, "sh_len_so", small_sh_kwargs(R, L, 60, 20))
    for R, L in ((60, 8), (61, 9), (60, 9), (61, 8)):
        n = f"sh_vel_so_{R}x{L}"
        yield n, lambda n=n, R=R, L=L: dump_sh(n, "sh_vel_so", small_sh_kwargs(R, L, 40, 20))
    for R, L in ((64, 8), (65, 9)):
        n = f"sh_len_adi_{R}x{L}"
        yield n, lambda n=n, R=R, L=L: dump_sh(n, "sh_len_adi", small_sh_kwargs(R, L, 40, 20))
    # all datastores (kernel 4)
    yield "sh_len_so_datastores_120x12", lambda: dump_sh(
        "sh_len_so_datastores_120x12", "sh_len_so", small_sh_kwargs(120, 12, 50, 20), extra_datastores=True
    )
    yield "sh_vel_so_datastores_120x12", lambda: dump_sh(
        "sh_vel_so_datastores_120x12", "sh_vel_so", small_sh_kwargs(120, 12, 50, 20), extra_datastores=True
    )

    # analysis at snapshot / data times
    yield "sh_len_so_analysis_90x8", lambda: dump_sh_analysis("sh_len_so_analysis_90x8", small_sh_kwargs(90, 8, 60, 20))

    # --- config 1: SH r_bound=100 a0, 500x50, Sinc 200 as + logistic window, 2000 steps
    def c1(kind):
        pw = 200 * u.asec
        rb = 100 * u.bohr_radius
        return dump_sh(
            f"c1_{kind}_500x50",
            kind,
            dict(
                r_bound=rb,
                r_points=500,
                l_bound=50,
                time_initial=-5 * pw,
                time_final=5 * pw,
                time_step=1 * u.asec,
                electric_potential=sinc(200),
                use_numeric_eigenstates=False,
                test_states=hydrogen_states(3),
                mask=ion.potentials.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb, smoothness=8),
            ),
            store=100,
        )

    yield "c1_sh_len_so_500x50", lambda: c1("sh_len_so")
    yield "c1_sh_vel_so_500x50", lambda: c1("sh_vel_so")

    # --- LineMesh small cases (even and odd z_points) ---
    def line_kwargs(Z, steps):
        well = ion.potentials.GaussianPotential(potential_extrema=-10 * u.eV, width=5 * u.bohr_radius)
        st = ion.states.GaussianWellState.from_potential(well, u.electron_mass)
        zb = 100 * u.bohr_radius
        return dict(
            z_bound=zb,
            z_points=Z,
            test_mass=u.electron_mass,
            internal_potential=well,
            initial_state=st,
            electric_potential=ion.potentials.SincPulse(pulse_width=100 * u.asec, fluence=0.1 * u.Jcm2, phase=0.3),
            time_initial=-steps / 2 * u.asec,
            time_final=steps / 2 * u.asec,
            time_step=1 * u.asec,
            mask=ion.potentials.RadialCosineMask(inner_radius=0.8 * zb, outer_radius=zb, smoothness=8),
        )

    for kind in ("line_len_cn", "line_len_so", "line_vel_so"):
        for Z in (1024, 1023):
            n = f"{kind}_{Z}"
            yield n, lambda n=n, kind=kind, Z=Z: dump_line(n, kind, line_kwargs(Z, 50))

    for kind in ("line_len_cn", "line_len_so"):
        n = f"{kind}_datastores_512"
        yield n, lambda n=n, kind=kind: dump_line_datastores(n, kind, line_kwargs(512, 40))

    if not big:
        return

    # --- the six known answers: dev/meshes/mesh_refactoring_helper.py:30-86, :204-251 ---
    def sh_golden(kind):
        pw = 100 * u.asec
        pulse = ion.potentials.GaussianPulse.from_number_of_cycles(pulse_width=pw, fluence=1 * u.Jcm2, phase=0, number_of_cycles=3)
        return dump_sh(
            f"known_{kind}_500x200",
            kind,
            dict(
                time_initial=-4 * pw,
                time_final=4 * pw,
                time_step=1 * u.asec,
                electric_potential=pulse,
                dc_correct_electric_potential=True,  # (sic) misspelt kwarg in the reference script: ignored
                r_bound=50 * u.bohr_radius,
                r_points=500,
                l_bound=200,
                theta_points=360,
                use_numeric_eigenstates=True,
                numeric_eigenstate_max_energy=20 * u.eV,
                numeric_eigenstate_max_angular_momentum=3,
            ),
            store=-1,
        )

    for kind in ("sh_len_so", "sh_len_adi", "sh_vel_so"):
        yield f"known_{kind}_500x200", lambda kind=kind: sh_golden(kind)

    def line_golden(kind):
        energy_spacing = 0.1 * u.eV
        test_mass = u.electron_mass
        qho = ion.potentials.HarmonicOscillator.from_energy_spacing_and_mass(energy_spacing=energy_spacing, mass=test_mass)
        line_states = [ion.states.QHOState.from_potential(qho, n=n, mass=test_mass) for n in range(5)]
        sine = ion.potentials.SineWave.from_photon_energy(photon_energy=energy_spacing, amplitude=0.0001 * u.atomic_electric_field)
        return dump_line(
            f"known_{kind}_4096",
            kind,
            dict(
                time_initial=0,
                time_final=1 * sine.period,
                time_step=0.001 * sine.period,
                internal_potential=qho,
                electric_potential=sine,
                initial_state=line_states[0],
                test_states=line_states,
                z_bound=100 * u.nm,
                z_points=2 ** 12,
            ),
            store=-1,
        )

    for kind in ("line_len_cn", "line_len_so", "line_vel_so"):
        yield f"known_{kind}_4096", lambda kind=kind: line_golden(kind)


KNOWN_ANSWERS = {  # dev/meshes/mesh_refactoring_helper.py:204-251
    "known_sh_len_adi_500x200": 0.312910470190,
    "known_sh_len_so_500x200": 0.312928752359,
    "known_sh_vel_so_500x200": 0.319513371899,
    "known_line_len_cn_4096": 0.370010185740,
    "known_line_len_so_4096": 0.370008474418,
    "known_line_vel_so_4096": 0.370924310122,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    for name, thunk in cases(args.big):
        if args.only is not None and args.only not in name:
            continue
        thunk()


if __name__ == "__main__":
    main()
