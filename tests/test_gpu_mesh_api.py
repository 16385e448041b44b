"""GPU tests of the reference-facing API: ``Specification(...).to_sim().run()`` with this package's own host-side
inputs must reproduce what the UNMODIFIED reference produced for the same physical parameters (tests/golden),
within 1e-10 (BASELINE.json north_star)."""
import numpy as np
import pytest

import ionization_b200 as ion
from ionization_b200 import potentials as P
from ionization_b200 import states as S
from ionization_b200 import units as u
from conftest import load_golden, rel_err
from test_host_layer import c1_spec

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _ips(sim):
    return np.array([sim.data.inner_products[s] for s in sim.spec.test_states]).T


def _small_len_spec(store=1, pulse=None, **kw):
    rb = 30 * u.bohr_radius
    pulse = pulse if pulse is not None else P.SincPulse(pulse_width=20 * u.asec, fluence=1 * u.Jcm2, phase=0)
    args = dict(electric_potential=pulse, mask=P.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb), time_initial=-30 * u.asec,
                time_final=30 * u.asec, r_points=100, l_bound=10, r_bound=rb, store_data_every=store)
    args.update(kw)
    return c1_spec("LEN", **args)


@pytest.mark.parametrize("gauge, fixture", [("LEN", "c1_sh_len_so_500x50"), ("VEL", "c1_sh_vel_so_500x50")])
def test_config1_run_matches_reference(gauge, fixture):
    ref = load_golden(fixture)
    sim = c1_spec(gauge).to_sim()
    sim.run()
    assert sim.status == ion.core.Status.FINISHED
    assert sim.time_index == sim.time_steps - 1 and sim.data_time_index == sim.data_time_steps
    assert np.max(np.abs(sim.data.norm - ref["norm"])) < TOL
    assert np.max(np.abs(_ips(sim) - ref["inner_products"])) < TOL
    assert rel_err(sim.mesh.g, ref["g_final"]) < TOL
    assert abs(sim.data.initial_state_overlap[-1] - float(ref["initial_state_overlap_final"])) < TOL
    ion_frac = 1 - sim.data.bound_state_overlap[-1]
    bound = ref["state_bound"].astype(bool)
    ion_ref = 1 - np.sum(np.abs(ref["inner_products"][-1][bound]) ** 2)
    assert abs(ion_frac - ion_ref) <= TOL * abs(ion_ref)
    e = sim.spec.electric_potential.get_electric_field_amplitude(sim.data_times)
    assert np.allclose(sim.data.electric_field_amplitude, e, rtol=1e-13, atol=0)


def test_callback_path_equals_device_resident_path():
    """run(callback=...) hands the live sim to user code after every step (mesh/sims.py:307); mesh.g must be current"""
    ref = load_golden("sh_len_so_100x10")
    a = _small_len_spec().to_sim()
    a.run()
    seen = []
    b = _small_len_spec().to_sim()
    b.run(callback=lambda s: seen.append((s.time_index, s.mesh.norm(), np.array(s.mesh.g, copy=True))))
    assert [t for t, _, _ in seen] == list(range(b.time_steps))
    assert rel_err(a.mesh.g, ref["g_final"]) < TOL and rel_err(b.mesh.g, ref["g_final"]) < TOL
    assert np.max(np.abs(a.data.norm - ref["norm"])) < TOL and np.max(np.abs(b.data.norm - ref["norm"])) < TOL
    assert np.max(np.abs(_ips(a) - ref["inner_products"])) < TOL
    assert np.max(np.abs(np.array([n for _, n, _ in seen]) - ref["norm"])) < TOL
    assert rel_err(seen[-1][2], ref["g_final"]) < TOL


def test_all_datastores_through_the_api():
    ref = load_golden("sh_len_so_datastores_120x12")
    D = ion.mesh
    radii = [r * u.bohr_radius for r in (5, 10, 20)]
    sim = _small_len_spec(
        time_initial=-25 * u.asec, time_final=25 * u.asec, r_points=120, l_bound=12,
        datastores=[D.Fields(), D.Norm(), D.InnerProducts(), D.InternalEnergyExpectationValue(), D.TotalEnergyExpectationValue(), D.ZExpectationValue(),
                    D.RExpectationValue(), D.NormWithinRadius(radii=radii), D.NormBySphericalHarmonic()],
    ).to_sim()
    sim.run()
    assert np.max(np.abs(sim.data.norm - ref["norm"])) < TOL
    assert rel_err(sim.data.internal_energy_expectation_value, ref["internal_energy"]) < TOL
    assert rel_err(sim.data.total_energy_expectation_value, ref["total_energy"]) < TOL
    assert rel_err(sim.data.r_expectation_value, ref["r_expectation"]) < TOL
    assert np.max(np.abs(sim.data.z_expectation_value - ref["z_expectation"])) < TOL * np.max(np.abs(ref["r_expectation"]))
    for k, r in enumerate(sorted(radii)):
        assert np.max(np.abs(sim.data.norm_within_radius[r] - ref["norm_within_radius"][:, k])) < TOL
    nbl = np.array([sim.data.norm_by_sph_harm[sh] for sh in sim.spec.spherical_harmonics]).T
    assert np.max(np.abs(nbl - ref["norm_by_l"])) < TOL
    assert rel_err(sim.data.electric_field_amplitude, ref["electric_field_amplitude"]) < 1e-12
    assert np.max(np.abs(sim.data.vector_potential_amplitude - ref["vector_potential_amplitude"])) < 1e-12 * np.max(np.abs(ref["vector_potential_amplitude"]))


@pytest.mark.parametrize("kind", ["line_len_so", "line_vel_so", "line_len_cn"])
@pytest.mark.parametrize("Z", [1024, 1023])
def test_line_simulation_matches_reference(kind, Z):
    ref = load_golden(f"{kind}_{Z}")
    well = P.GaussianPotential(potential_extrema=-10 * u.eV, width=5 * u.bohr_radius)
    zb = 100 * u.bohr_radius
    ops = ion.mesh.LineVelocityGaugeOperators() if kind == "line_vel_so" else ion.mesh.LineLengthGaugeOperators()
    sim = ion.mesh.LineSpecification(
        "line", z_bound=zb, z_points=Z, test_mass=u.electron_mass, internal_potential=well, initial_state=S.GaussianWellState.from_potential(well, u.electron_mass),
        electric_potential=P.SincPulse(pulse_width=100 * u.asec, fluence=0.1 * u.Jcm2, phase=0.3), time_initial=-25 * u.asec, time_final=25 * u.asec,
        time_step=1 * u.asec, mask=P.RadialCosineMask(inner_radius=0.8 * zb, outer_radius=zb, smoothness=8), operators=ops,
        evolution_method=ion.mesh.AlternatingDirectionImplicit() if kind == "line_len_cn" else ion.mesh.SplitInteractionOperator(),
    ).to_sim()
    sim.run()
    assert rel_err(sim.mesh.g, ref["g_final"]) < TOL
    assert np.max(np.abs(sim.data.norm - ref["norm"])) < TOL
    assert np.max(np.abs(_ips(sim) - ref["inner_products"])) < TOL


def test_known_answer_through_the_api_with_numeric_eigenstates():
    """dev/meshes/mesh_refactoring_helper.py:30-86,:204-251 -- LEN SO: final initial-state overlap 0.312928752359"""
    pw = 100 * u.asec
    sim = ion.mesh.SphericalHarmonicSpecification(
        "known", time_initial=-4 * pw, time_final=4 * pw, time_step=1 * u.asec,
        electric_potential=P.GaussianPulse.from_number_of_cycles(pulse_width=pw, fluence=1 * u.Jcm2, phase=0, number_of_cycles=3),
        r_bound=50 * u.bohr_radius, r_points=500, l_bound=200, theta_points=360, use_numeric_eigenstates=True,
        numeric_eigenstate_max_energy=20 * u.eV, numeric_eigenstate_max_angular_momentum=3, store_data_every=-1,
        operators=ion.mesh.SphericalHarmonicLengthGaugeOperators(), evolution_method=ion.mesh.SplitInteractionOperator(),
    ).to_sim()
    sim.run()
    assert abs(sim.data.initial_state_overlap[-1] - 0.312928752359) < 1e-10
    assert abs(sim.data.norm[-1] - 1.0) < 1e-10


def test_field_free_evolution_keeps_norm_and_overlaps():
    """tests/mesh/test_sims.py:31-82 of the reference (atol 1e-14 there; 1e-13 here)"""
    for ops in (ion.mesh.SphericalHarmonicLengthGaugeOperators(), ion.mesh.SphericalHarmonicVelocityGaugeOperators()):
        for n, l in ((1, 0), (2, 0), (2, 1)):
            sim = ion.mesh.SphericalHarmonicSpecification(
                "test", initial_state=S.HydrogenBoundState(n, l), operators=ops, evolution_method=ion.mesh.SplitInteractionOperator(), time_initial=0,
                time_final=100 * u.asec, time_step=1 * u.asec, r_bound=50 * u.bohr_radius, r_points=250, l_bound=30, use_numeric_eigenstates=True,
                numeric_eigenstate_max_energy=10 * u.eV, numeric_eigenstate_max_angular_momentum=10,
            ).to_sim()
            sim.run()
            ov = np.array([v for v in sim.data.state_overlaps.values()])
            assert abs(sim.data.norm[0] - sim.data.norm[-1]) < 1e-13
            assert np.max(np.abs(ov[:, 0] - ov[:, -1])) < 1e-13


def test_checkpoint_and_resume(tmp_path):
    """mesh/sims.py:327-356, :405-438: a pickled, partly evolved simulation resumes to the same answer"""
    ref = load_golden("sh_len_so_100x10")
    sim = _small_len_spec().to_sim()

    class Stop(Exception):
        pass

    def cb(s):
        if s.time_index == 25:
            s.save(tmp_path)
            raise Stop

    with pytest.raises(Stop):
        sim.run(callback=cb)
    resumed = ion.mesh.MeshSimulation.load(tmp_path / "c1.sim")
    assert resumed.time_index == 25
    resumed.run()
    assert rel_err(resumed.mesh.g, ref["g_final"]) < TOL
    assert np.max(np.abs(resumed.data.norm - ref["norm"])) < TOL


def test_ensemble_equals_individual_runs():
    """scan ensemble = cartesian product of pulse parameters over one mesh (ionization_scans/scan_mesh.py:40-68)"""
    pulses = [P.SincPulse(pulse_width=20 * u.asec, fluence=flu * u.Jcm2, phase=ph) for flu in (0.1, 1.0) for ph in (0.0, 1.0, 2.5)]
    specs = [_small_len_spec(store=10, pulse=p) for p in pulses]
    sims = ion.mesh.run_ensemble(specs)
    ref = load_golden("sh_len_so_100x10")
    assert rel_err(sims[3].mesh.g, ref["g_final"]) < TOL  # (fluence 1, phase 0) is the fixture's pulse
    for p, s in zip(pulses, sims):
        single = _small_len_spec(store=10, pulse=p).to_sim()
        single.run()
        assert rel_err(s.mesh.g, single.mesh.g) < 1e-13
        assert np.max(np.abs(s.data.norm - single.data.norm)) < 1e-13
        assert np.max(np.abs(_ips(s) - _ips(single))) < 1e-13


def test_spherical_harmonic_adi_through_the_api():
    """SphericalHarmonicLengthGaugeOperators + AlternatingDirectionImplicit (evolution_methods.py:46-77):
    the spec of the reference fixture sh_len_adi_64x8 (oracle/make_golden.py: small_sh_kwargs(64, 8, 40, 20))."""
    ref = load_golden("sh_len_adi_64x8")
    rb = 30 * u.bohr_radius
    spec = c1_spec(
        "LEN", r_bound=rb, r_points=64, l_bound=8, time_initial=-20 * u.asec, time_final=20 * u.asec,
        electric_potential=P.SincPulse(pulse_width=20 * u.asec, fluence=1 * u.Jcm2, phase=0),
        mask=P.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb, smoothness=8),
        evolution_method=ion.mesh.AlternatingDirectionImplicit(), store_data_every=1,
    )
    sim = spec.to_sim()
    sim.run()
    assert np.max(np.abs(sim.data.norm - ref["norm"])) < TOL
    assert np.max(np.abs(_ips(sim) - ref["inner_products"])) < TOL
    assert rel_err(sim.mesh.g, ref["g_final"]) < TOL
