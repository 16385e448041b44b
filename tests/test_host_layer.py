"""CPU tests of the host-side mirror of ``ionization.mesh``: the inputs it hands to the CUDA engine must equal the ones
the UNMODIFIED reference builds (fixtures dumped from live reference objects, oracle/make_golden.py), the API surface
(datastore names, exceptions, pickling) must behave like the reference's (tests/mesh/test_datastores.py,
test_save_and_load.py), and nothing may silently compute on the CPU when no GPU is present."""
import pickle

import numpy as np
import pytest

import ionization_b200 as ion
from ionization_b200 import coefficients as C
from ionization_b200 import potentials as P
from ionization_b200 import states as S
from ionization_b200 import units as u
from conftest import load_golden, rel_err


def c1_spec(gauge="LEN", **kw):
    pw = 200 * u.asec
    rb = 100 * u.bohr_radius
    ops = ion.mesh.SphericalHarmonicLengthGaugeOperators() if gauge == "LEN" else ion.mesh.SphericalHarmonicVelocityGaugeOperators()
    args = dict(
        r_bound=rb, r_points=500, l_bound=50, time_initial=-5 * pw, time_final=5 * pw, time_step=1 * u.asec,
        electric_potential=P.SincPulse(pulse_width=pw, fluence=1 * u.Jcm2, phase=0, window=P.LogisticWindow(window_time=4 * pw, window_width=0.2 * pw)),
        use_numeric_eigenstates=False, test_states=[S.HydrogenBoundState(n, l) for n in range(1, 4) for l in range(n)],
        mask=P.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb, smoothness=8), operators=ops,
        evolution_method=ion.mesh.SplitInteractionOperator(), store_data_every=100,
    )
    args.update(kw)
    return ion.mesh.SphericalHarmonicSpecification("c1", **args)


@pytest.mark.parametrize("gauge, fixture", [("LEN", "c1_sh_len_so_500x50"), ("VEL", "c1_sh_vel_so_500x50")])
def test_spherical_harmonic_inputs_equal_the_reference(gauge, fixture):
    ref = load_golden(fixture)
    sim = c1_spec(gauge).to_sim()
    assert np.array_equal(sim.times, ref["times"])
    assert np.array_equal(sim._taus, ref["taus"])
    assert rel_err(sim._fields, ref["fields"]) < 1e-13
    assert np.array_equal(sim.data_times, ref["data_times"])
    assert np.array_equal(sim._mask_vector, ref["mask"])
    assert rel_err(sim.mesh.g, ref["g0"]) < 1e-14
    hd, ho = sim.spec.operators.hamiltonian_vectors(sim.mesh)
    assert rel_err(hd, ref["h_diag"]) < 1e-14 and rel_err(ho, ref["h_off"]) < 1e-15
    assert [s.l for s in sim._flat_states] == list(ref["state_l"])
    rows = np.array([sim.mesh.get_radial_g_for_state(s) for s in sim._flat_states])
    assert rel_err(rows, ref["state_rows"]) < 1e-14
    if gauge == "LEN":
        c_l, x_j = C.sh_len_coupling(sim.mesh.r, 50, sim.spec.test_charge)
        assert rel_err(c_l, ref["c_l"]) < 1e-15 and rel_err(x_j, ref["x_j"]) < 1e-15
    else:
        c_l, f1, y, z = C.sh_vel_coupling(sim.mesh.r, sim.mesh.delta_r, 50, sim.spec.test_charge, sim.spec.test_mass)
        for a, k in ((c_l, "c_l"), (f1, "f1_l"), (y, "y_j"), (z, "z_j")):
            assert rel_err(a, ref[k]) < 1e-15


@pytest.mark.parametrize("kind", ["line_len_cn", "line_len_so", "line_vel_so"])
@pytest.mark.parametrize("Z", [1024, 1023])
def test_line_inputs_equal_the_reference(kind, Z):
    ref = load_golden(f"{kind}_{Z}")
    well = P.GaussianPotential(potential_extrema=-10 * u.eV, width=5 * u.bohr_radius)
    zb = 100 * u.bohr_radius
    ops = ion.mesh.LineVelocityGaugeOperators() if kind == "line_vel_so" else ion.mesh.LineLengthGaugeOperators()
    method = ion.mesh.AlternatingDirectionImplicit() if kind == "line_len_cn" else ion.mesh.SplitInteractionOperator()
    sim = ion.mesh.LineSpecification(
        "line", z_bound=zb, z_points=Z, test_mass=u.electron_mass, internal_potential=well, initial_state=S.GaussianWellState.from_potential(well, u.electron_mass),
        electric_potential=P.SincPulse(pulse_width=100 * u.asec, fluence=0.1 * u.Jcm2, phase=0.3), time_initial=-25 * u.asec, time_final=25 * u.asec,
        time_step=1 * u.asec, mask=P.RadialCosineMask(inner_radius=0.8 * zb, outer_radius=zb, smoothness=8), operators=ops, evolution_method=method,
    ).to_sim()
    assert sim._program == kind
    assert np.array_equal(sim.times, ref["times"])
    assert rel_err(sim._fields, ref["fields"]) < 1e-13
    assert rel_err(sim.mesh.g, ref["g0"]) < 1e-14
    hd, ho = sim.spec.operators.hamiltonian_vectors(sim.mesh)
    assert rel_err(hd[0], ref["h_diag"]) < 1e-14 and rel_err(ho, ref["h_off"]) < 1e-15
    w_z, v_pref = C.line_coupling(sim.mesh.z_mesh, sim.mesh.delta_z, sim.spec.test_charge, sim.spec.test_mass)
    assert rel_err(w_z, ref["w_z"]) < 1e-15 and abs(v_pref - float(ref["v_pref"])) <= 1e-15 * abs(v_pref)
    assert np.array_equal(sim._mask_vector, ref["mask"])


def test_numeric_eigenstates_match_the_reference_basis():
    """SphericalHarmonicMesh.get_numeric_eigenstate_basis (meshes.py:1281-1356): same states (energy <= 20 eV, l <= 3),
    same radial vectors up to the arbitrary phase of an eigenvector."""
    ref = load_golden("known_sh_len_so_500x200")
    pw = 100 * u.asec
    sim = ion.mesh.SphericalHarmonicSpecification(
        "known", time_initial=-4 * pw, time_final=4 * pw, time_step=1 * u.asec,
        electric_potential=P.GaussianPulse.from_number_of_cycles(pulse_width=pw, fluence=1 * u.Jcm2, phase=0, number_of_cycles=3),
        r_bound=50 * u.bohr_radius, r_points=500, l_bound=200, use_numeric_eigenstates=True, numeric_eigenstate_max_energy=20 * u.eV,
        numeric_eigenstate_max_angular_momentum=3, store_data_every=-1,
    ).to_sim()
    assert len(sim.spec.test_states) == len(ref["state_l"]) == 77
    assert [s.l for s in sim._flat_states] == list(ref["state_l"])
    rows = np.array([sim.mesh.get_radial_g_for_state(s) for s in sim._flat_states])
    for a, b in zip(rows, ref["state_rows"]):
        # ARPACK returns complex eigenvectors with an arbitrary phase for the reference's complex128 matrices
        phase = np.vdot(a, b) / abs(np.vdot(a, b))
        assert rel_err(a * phase, b) < 1e-8
    assert np.allclose(np.abs(sim.mesh.g), np.abs(ref["g0"]), atol=1e-8 * np.max(np.abs(ref["g0"])))
    assert rel_err(sim._fields, ref["fields"]) < 1e-13
    assert sim.mesh.norm(sim.mesh.g) == pytest.approx(1.0, abs=1e-12)  # tests/mesh/test_sims.py:12-29 of the reference


def test_vector_potential_prefix_rule_equals_the_per_step_definition():
    pulse = P.SincPulse(pulse_width=100 * u.asec, fluence=1 * u.Jcm2, phase=0.5)
    times = np.linspace(-300 * u.asec, 300 * u.asec, 601)
    fast = C.vector_potential_series(pulse, times)
    slow = np.array([pulse.get_vector_potential_amplitude_numeric(times[: n + 1]) for n in range(1, len(times))])
    assert rel_err(fast, slow) < 1e-13
    # non-uniform grid
    t2 = np.cumsum(np.concatenate([[0], np.random.default_rng(0).uniform(0.5, 1.5, 200)])) * u.asec
    assert rel_err(C.vector_potential_series(pulse, t2), np.array([pulse.get_vector_potential_amplitude_numeric(t2[: n + 1]) for n in range(1, len(t2))])) < 1e-13


def test_dc_correction_zeroes_the_vector_potential():
    """tests/potentials/test_pulse_dc_correction.py of the reference"""
    pw = 100 * u.asec
    times = np.linspace(-10 * pw, 10 * pw, 2001)
    pulse = P.SincPulse(pulse_width=pw, fluence=1 * u.Jcm2, phase=0)
    corrected = P.DC_correct_electric_potential(pulse, times)
    a0 = pulse.get_vector_potential_amplitude_numeric_cumulative(times)[-1]
    a1 = corrected.get_vector_potential_amplitude_numeric_cumulative(times)[-1]
    assert abs(a1) < 1e-6 * abs(a0)


# ---- masks: tests/potentials/test_masks.py of the reference ----------------------------------------------------
def test_radial_cosine_mask_end_points_and_validation():
    m = P.RadialCosineMask(inner_radius=10 * u.bohr_radius, outer_radius=20 * u.bohr_radius, smoothness=8)
    assert m(r=10 * u.bohr_radius) == 1
    assert np.allclose(m(r=20 * u.bohr_radius), 0, atol=1e-14)
    assert m(r=5 * u.bohr_radius) == 1 and m(r=25 * u.bohr_radius) == 0
    with pytest.raises(ion.exceptions.InvalidMaskParameter):
        P.RadialCosineMask(inner_radius=-1, outer_radius=1)
    with pytest.raises(ion.exceptions.InvalidMaskParameter):
        P.RadialCosineMask(inner_radius=2, outer_radius=1)
    with pytest.raises(ion.exceptions.InvalidMaskParameter):
        P.RadialCosineMask(inner_radius=1, outer_radius=2, smoothness=0.5)
    from oracle import restate

    r = np.linspace(0, 30, 301) * u.bohr_radius
    assert np.array_equal(m(r=r), restate.radial_cosine_mask(r, m.inner_radius, m.outer_radius, m.smoothness))


# ---- datastores: tests/mesh/test_datastores.py of the reference ----------------------------------------------------
DATASTORE_TYPES = [
    ion.mesh.Fields, ion.mesh.Norm, ion.mesh.InnerProducts, ion.mesh.InternalEnergyExpectationValue, ion.mesh.TotalEnergyExpectationValue,
    ion.mesh.ZExpectationValue, ion.mesh.RExpectationValue, ion.mesh.NormWithinRadius,
]


def _small_spec(spec_type, **kw):
    if spec_type is ion.mesh.LineSpecification:
        return spec_type("t", z_points=64, time_final=5 * u.asec, **kw)
    return spec_type("t", r_points=40, l_bound=4, r_bound=20 * u.bohr_radius, use_numeric_eigenstates=False, time_final=5 * u.asec, **kw)


@pytest.mark.parametrize("spec_type", [ion.mesh.LineSpecification, ion.mesh.SphericalHarmonicSpecification])
@pytest.mark.parametrize("ds_type", DATASTORE_TYPES)
def test_datastore_names_exist_and_missing_ones_raise(spec_type, ds_type):
    sim = _small_spec(spec_type, datastores=[ds_type()]).to_sim()
    for name in ion.mesh.DATASTORE_TYPE_TO_DATA_NAMES[ds_type]:
        getattr(sim.data, name)
    other = next(t for t in DATASTORE_TYPES if t is not ds_type)
    for name in ion.mesh.DATASTORE_TYPE_TO_DATA_NAMES[other]:
        with pytest.raises(ion.exceptions.MissingDatastore):
            getattr(sim.data, name)
    with pytest.raises(ion.exceptions.UnknownData):
        sim.data.no_such_data


def test_duplicate_datastores_rejected_and_to_sim_gives_fresh_datastores():
    with pytest.raises(ion.exceptions.DuplicateDatastores):
        _small_spec(ion.mesh.SphericalHarmonicSpecification, datastores=[ion.mesh.Norm(), ion.mesh.Norm()])
    spec = _small_spec(ion.mesh.SphericalHarmonicSpecification)
    a, b = spec.to_sim(), spec.to_sim()
    assert a.datastores_by_type[ion.mesh.Norm] is not b.datastores_by_type[ion.mesh.Norm]
    assert np.all(np.isnan(a.data.norm))


# ---- save / load: tests/mesh/test_save_and_load.py of the reference ----------------------------------------------------
@pytest.mark.parametrize("spec_type", [ion.mesh.LineSpecification, ion.mesh.SphericalHarmonicSpecification])
def test_saved_simulation_round_trips_the_mesh(spec_type, tmp_path):
    sim = _small_spec(spec_type).to_sim()
    special_g = np.random.default_rng(0).random(sim.mesh.g.shape) + 0j
    sim.mesh.g = special_g.copy()
    path = sim.save(tmp_path)
    loaded = ion.mesh.MeshSimulation.load(path)
    assert loaded.mesh is not sim.mesh
    assert np.array_equal(loaded.mesh.g, special_g)
    path = sim.save(tmp_path, save_mesh=False)
    assert ion.mesh.MeshSimulation.load(path).mesh is None
    assert sim.mesh is not None


def test_unsupported_combinations_fail_loudly():
    with pytest.raises(ion.exceptions.UnsupportedConfiguration):
        ion.mesh.SphericalHarmonicSpecification(
            "x", operators=ion.mesh.SphericalHarmonicVelocityGaugeOperators(), evolution_method=ion.mesh.AlternatingDirectionImplicit(),
            r_points=40, l_bound=4, use_numeric_eigenstates=False,
        ).to_sim()


def test_no_cpu_fallback_without_a_gpu():
    """the product path must fail loudly when there is no CUDA device (or no built extension)"""
    from ionization_b200 import engine

    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    sim = _small_spec(ion.mesh.SphericalHarmonicSpecification).to_sim()
    with pytest.raises(ion.exceptions.EngineError):
        sim.run()
    with pytest.raises(ion.exceptions.EngineError):
        engine.tdma((np.ones(3, complex), np.ones(4, complex) * 3, np.ones(3, complex)), np.ones(4, complex))
    # state pickles with the simulation
    pickle.dumps(sim)
