"""SURVEY 8(f)-4 / ADVICE r01: analysis evaluated at snapshot and data times -- plane-wave overlaps (snapshots.py:37-80,
meshes.py:1138-1190), the directional radial probability current (data.py:467-531, meshes.py:1358-1370) and the LineMesh
expectation-value datastores -- against fixtures produced by the reference's own objects (oracle/make_golden.py).
The CPU tests exercise the host-side analysis on the reference's final wavefunction; the GPU tests the whole run."""
import numpy as np
import pytest

import ionization_b200 as ion
from ionization_b200 import potentials as P
from ionization_b200 import states as S
from ionization_b200 import units as u
from conftest import load_golden, rel_err

TOL = 1e-10


def analysis_spec(**kw):
    rb = 30 * u.bohr_radius
    args = dict(
        r_bound=rb, r_points=90, l_bound=8, time_initial=-30 * u.asec, time_final=30 * u.asec, time_step=1 * u.asec,
        electric_potential=P.SincPulse(pulse_width=20 * u.asec, fluence=1 * u.Jcm2, phase=0), use_numeric_eigenstates=False,
        test_states=[S.HydrogenBoundState(n, l) for n in range(1, 4) for l in range(n)],
        mask=P.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb, smoothness=8), operators=ion.mesh.SphericalHarmonicLengthGaugeOperators(),
        evolution_method=ion.mesh.SplitInteractionOperator(), store_data_every=20, theta_points=24,
    )
    args.update(kw)
    return ion.mesh.SphericalHarmonicSpecification("analysis", **args)


def test_plane_wave_overlaps_and_radial_current_of_the_reference_wavefunction():
    ref = load_golden("sh_len_so_analysis_90x8")
    sim = analysis_spec().to_sim()
    sim.mesh.g = ref["g_final"]
    thetas, ks = ref["snapshot_thetas"], ref["snapshot_wavenumbers"]
    th, kk, ip = sim.mesh.inner_product_with_plane_waves(thetas, ks)
    assert th.shape == kk.shape == ip.shape == (len(thetas), len(ks))
    assert rel_err(ip, ref["snapshot_60_inner_product_with_plane_waves"]) < TOL
    g_free = sim.mesh.get_g_with_states_removed(sim.bound_states, sim.mesh.g)
    assert rel_err(sim.mesh.inner_product_with_plane_waves(thetas, ks, g=g_free)[2], ref["snapshot_60_inner_product_with_plane_waves__free_only"]) < TOL
    assert rel_err(sim.mesh.get_radial_probability_current_density_mesh__spatial(), ref["radial_current_density_final"]) < TOL


def test_snapshot_times_by_index_and_by_time_and_datastore_registry():
    sim = analysis_spec(snapshot_indices=(20,), snapshot_times=(10.2 * u.asec,)).to_sim()
    assert sorted(sim._snapshot_indices) == [20, 40]
    assert sim.snapshot_times == {sim.times[20], sim.times[40]}
    assert ion.mesh.DATA_NAME_TO_DATASTORE_TYPE["radial_probability_current__total"] is ion.mesh.DirectionalRadialProbabilityCurrent
    assert "radial_probability_current__pos_z" in ion.mesh.DATASTORE_TYPE_TO_DATA_NAMES[ion.mesh.DirectionalRadialProbabilityCurrent]
    with pytest.raises(ion.exceptions.MissingDatastore):
        sim.data.radial_probability_current__pos_z


@pytest.mark.gpu
def test_snapshots_and_radial_current_through_a_run():
    ref = load_golden("sh_len_so_analysis_90x8")
    D = ion.mesh
    snap_kw = dict(plane_wave_overlap__max_wavenumber=float(ref["snapshot_kwargs_max_wavenumber"]), plane_wave_overlap__wavenumber_points=12,
                   plane_wave_overlap__theta_points=9)
    sim = analysis_spec(
        snapshot_indices=(20, 60), snapshot_type=D.SphericalHarmonicSnapshot, snapshot_kwargs=snap_kw,
        datastores=[D.Fields(), D.Norm(), D.InnerProducts(), D.DirectionalRadialProbabilityCurrent()],
    ).to_sim()
    sim.run()
    assert rel_err(sim.mesh.g, ref["g_final"]) < TOL
    assert np.max(np.abs(sim.data.norm - ref["norm"])) < TOL
    assert sorted(sim.snapshots) == [20, 60]
    for idx in (20, 60):
        snap = sim.snapshots[idx]
        assert abs(snap.data["norm"] - float(ref[f"snapshot_{idx}_norm"])) < TOL
        for key in ("inner_product_with_plane_waves", "inner_product_with_plane_waves__free_only"):
            th, kk, ip = snap.data[key]
            assert np.allclose(th[:, 0], ref["snapshot_thetas"]) and np.allclose(kk[0], ref["snapshot_wavenumbers"])
            assert rel_err(ip, ref[f"snapshot_{idx}_{key}"]) < TOL, (idx, key)
    scale = np.max(np.abs(ref["radial_current_pos_z_final"]))
    assert np.max(np.abs(sim.data.radial_probability_current__pos_z[-1] - ref["radial_current_pos_z_final"])) < TOL * scale
    assert np.max(np.abs(sim.data.radial_probability_current__neg_z[-1] - ref["radial_current_neg_z_final"])) < TOL * scale
    assert np.allclose(sim.data.radial_probability_current__total[-1], ref["radial_current_pos_z_final"] + ref["radial_current_neg_z_final"], rtol=0, atol=TOL * scale)
    assert not np.any(np.isnan(sim.data.radial_probability_current__pos_z))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["line_len_cn", "line_len_so"])
def test_line_mesh_expectation_value_datastores(kind):
    """<z> and the energies on a LineMesh (the engine's "r" observable stands in for <z> there)"""
    ref = load_golden(f"{kind}_datastores_512")
    D = ion.mesh
    well = P.GaussianPotential(potential_extrema=-10 * u.eV, width=5 * u.bohr_radius)
    zb = 100 * u.bohr_radius
    method = D.AlternatingDirectionImplicit() if kind == "line_len_cn" else D.SplitInteractionOperator()
    sim = D.LineSpecification(
        "line_ds", z_bound=zb, z_points=512, test_mass=u.electron_mass, internal_potential=well, initial_state=S.GaussianWellState.from_potential(well, u.electron_mass),
        electric_potential=P.SincPulse(pulse_width=100 * u.asec, fluence=0.1 * u.Jcm2, phase=0.3), time_initial=-20 * u.asec, time_final=20 * u.asec,
        time_step=1 * u.asec, mask=P.RadialCosineMask(inner_radius=0.8 * zb, outer_radius=zb, smoothness=8), operators=D.LineLengthGaugeOperators(),
        evolution_method=method, store_data_every=5,
        datastores=[D.Fields(), D.Norm(), D.InnerProducts(), D.InternalEnergyExpectationValue(), D.TotalEnergyExpectationValue(), D.ZExpectationValue()],
    ).to_sim()
    sim.run()
    assert rel_err(sim.mesh.g, ref["g_final"]) < TOL
    assert np.max(np.abs(sim.data.norm - ref["norm"])) < TOL
    assert rel_err(sim.data.internal_energy_expectation_value, ref["internal_energy"]) < TOL
    assert np.max(np.abs(sim.data.z_expectation_value - ref["z_expectation"])) < TOL * float(zb)
    # total energy = <H0> + E(t) (-q) <z> on a line (mesh_operators.py:320-327; the reference itself cannot evaluate it there)
    e = sim.spec.electric_potential.get_electric_field_amplitude(sim.data_times)
    expect = ref["internal_energy"] + e * (-sim.spec.test_charge) * ref["z_expectation"]
    assert rel_err(sim.data.total_energy_expectation_value, expect) < 1e-9
    assert abs(sim.mesh.z_expectation_value() - ref["z_expectation"][-1]) < TOL * float(zb)
