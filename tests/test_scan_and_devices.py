"""Scan container / runner (ionization/analysis.py:65-123, ionization_scans/scan_utils.py:638-663) and the multi-GPU entry points
of the mesh API: run_ensemble(specs, devices=[...]) and SphericalHarmonicSpecification(..., devices=[...]) (l-block shards)."""
import gzip
import pickle

import numpy as np
import pytest

import ionization_b200 as ion
from ionization_b200 import potentials as P
from ionization_b200 import scan
from ionization_b200 import units as u
from conftest import rel_err
from test_host_layer import c1_spec

TOL = 1e-10


def _specs(n, **kw):
    rb = 30 * u.bohr_radius
    out = []
    for i in range(n):
        args = dict(
            electric_potential=P.SincPulse(pulse_width=20 * u.asec, fluence=(0.5 + 0.3 * i) * u.Jcm2, phase=0.4 * i), time_initial=-20 * u.asec, time_final=20 * u.asec,
            mask=P.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb), r_points=80, l_bound=8, r_bound=rb, store_data_every=10,
        )
        args.update(kw)
        spec = c1_spec("LEN", **args)
        spec.name = spec.file_name = f"member_{i}"
        spec.fluence_index = i
        out.append(spec)
    return out


class _FakeSpec:
    def __init__(self, k):
        self.k, self.parity = k, k % 2


class _FakeSim:
    def __init__(self, k):
        self.spec = _FakeSpec(k)


def test_parameter_scan_round_trip_and_selection(tmp_path):
    ps = scan.ParameterScan("tag", [_FakeSim(k) for k in range(5)])
    path = ps.save(tmp_path)
    assert path.name == "tag.sims"
    with gzip.open(path, "rb") as f:  # export_scan.py:39-47: count first, then one pickle per simulation
        assert pickle.load(f) == 5
    back = scan.ParameterScan.from_file(path)
    assert len(back) == 5 and back.tag == "tag" and [s.spec.k for s in back] == list(range(5))
    assert back.parameter_set("parity") == {0, 1}
    assert [s.spec.k for s in back.select(parity=1)] == [1, 3]
    assert back[2].spec.k == 2
    # analysis.py:84-85: a single pickled list is accepted as well
    with gzip.open(tmp_path / "list.sims", "wb") as f:
        pickle.dump([_FakeSim(7)], f)
    assert scan.ParameterScan.from_file(tmp_path / "list.sims")[0].spec.k == 7


def test_ensemble_members_get_their_own_dc_correction():
    """every member of an ensemble gets the corrections MeshSimulation.__init__ applies (mesh/sims.py:53-75)"""
    from ionization_b200.mesh import ensemble

    specs = _specs(3, electric_potential_dc_correction=True)
    singles = [s.to_sim()._fields for s in _specs(3, electric_potential_dc_correction=True)]
    ens = ensemble.MeshEnsemble(specs, device=0)
    for sim, ref in zip(ens.sims, singles):
        assert np.array_equal(sim._fields, ref)
    assert not np.array_equal(ens.sims[0]._fields, ens.sims[1]._fields)
    uncorrected = _specs(2)[1].to_sim()._fields
    assert not np.array_equal(ens.sims[1]._fields, uncorrected)


def test_ensemble_rejects_members_with_another_hamiltonian_or_other_datastores():
    from ionization_b200.mesh import ensemble

    a, b = _specs(2)
    b.internal_potential = P.CoulombPotential(charge=2 * u.proton_charge)
    with pytest.raises(ion.exceptions.UnsupportedConfiguration):
        ensemble.MeshEnsemble([a, b])
    a, b = _specs(2)
    b.datastores = b.datastores[:-1]
    b.datastore_types = tuple(sorted(set(ds.__class__ for ds in b.datastores), key=lambda ds: ds.__name__))
    with pytest.raises(ion.exceptions.UnsupportedConfiguration):
        ensemble.MeshEnsemble([a, b])


@pytest.mark.gpu
def test_scan_over_two_device_blocks_equals_member_by_member_runs(tmp_path):
    """run_scan splits the members into contiguous blocks, one batched device run per entry of `devices` (here twice GPU 0)"""
    singles = [s.to_sim().run() for s in _specs(5)]
    sims = scan.run_scan(_specs(5), devices=[0, 0])
    assert [s.name for s in sims] == [f"member_{i}" for i in range(5)]
    for a, b in zip(sims, singles):
        assert a.mesh is None  # scan_utils.run strips the mesh (:656-661)
        assert np.max(np.abs(a.data.norm - b.data.norm)) < TOL
        assert abs(a.data.initial_state_overlap[-1] - b.data.initial_state_overlap[-1]) < TOL
    kept = ion.mesh.run_ensemble(_specs(4), devices=[0, 0])
    for a, b in zip(kept, singles):
        assert rel_err(a.mesh.g, b.mesh.g) < TOL
    path = scan.ParameterScan("scan", sims).save(tmp_path)
    back = scan.ParameterScan.from_file(path)
    assert len(back) == 5 and np.array_equal(back[3].data.norm, sims[3].data.norm)
    assert [s.name for s in back.select(fluence_index=2)] == ["member_2"]


@pytest.mark.gpu
@pytest.mark.parametrize("gauge", ["LEN", "VEL"])
def test_one_simulation_l_block_sharded_behind_the_mesh_api(gauge):
    """SphericalHarmonicSpecification(devices=[...]).to_sim().run(): shards linked by the engine's peer-memory halo exchange (here
    three shards on GPU 0, one host thread each), datastores from combined partial observations"""
    D = ion.mesh
    rb = 30 * u.bohr_radius

    def spec(**kw):
        args = dict(
            electric_potential=P.SincPulse(pulse_width=20 * u.asec, fluence=5 * u.Jcm2, phase=0), time_initial=-30 * u.asec, time_final=30 * u.asec,
            mask=P.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb), r_points=100, l_bound=12, r_bound=rb, store_data_every=7,
            datastores=[D.Fields(), D.Norm(), D.InnerProducts(), D.NormBySphericalHarmonic(), D.RExpectationValue(), D.NormWithinRadius(radii=[5 * u.bohr_radius, 12 * u.bohr_radius])],
        )
        args.update(kw)
        return c1_spec(gauge, **args)

    ref = spec().to_sim().run()
    sim = spec(devices=[0, 0, 0]).to_sim().run()
    assert rel_err(sim.mesh.g, ref.mesh.g) < 1e-12
    assert np.max(np.abs(sim.data.norm - ref.data.norm)) < 1e-12
    for s in ref.spec.test_states:
        assert np.max(np.abs(sim.data.inner_products[s] - ref.data.inner_products[s])) < 1e-12
    assert rel_err(sim.data.r_expectation_value, ref.data.r_expectation_value) < 1e-12
    for r in ref.data.norm_within_radius:
        assert np.max(np.abs(sim.data.norm_within_radius[r] - ref.data.norm_within_radius[r])) < 1e-12
    for sh in ref.data.norm_by_l:
        assert np.max(np.abs(sim.data.norm_by_l[sh] - ref.data.norm_by_l[sh])) < 1e-12
    with pytest.raises(ion.exceptions.UnsupportedConfiguration):
        spec(devices=[0, 0], datastores=[D.Norm(), D.ZExpectationValue()]).to_sim().run()


@pytest.mark.gpu
@pytest.mark.parametrize("program", ["sh_len_so", "line_len_cn", "sh_vel_so"])
def test_device_field_setup_equals_the_host_path(program):
    """E(t) / A(t) of a whole scan of windowed Sinc pulses in two kernels (csrc/fields.cuh) against the host builders,
    which are checked against the reference's own per-step values (tests/test_host_layer.py)"""
    from ionization_b200 import coefficients as C

    pw = 200 * u.asec
    times = C.time_grid(-5 * pw, 5 * pw, 1 * u.asec)
    pulses = [P.SincPulse(pulse_width=pw, fluence=f * u.Jcm2, phase=ph, window=P.LogisticWindow(window_time=4 * pw, window_width=0.2 * pw))
              for f in np.geomspace(0.01, 20, 7) for ph in np.linspace(0, u.twopi, 5, endpoint=False)]
    pulses.append(P.SincPulse(pulse_width=93 * u.asec, fluence=2 * u.Jcm2, phase=1.0, pulse_center=50 * u.asec))  # no window, off-centre
    host = C.field_series_batch(program, pulses, times, 1 * u.asec, device=None)
    dev = C.field_series_batch(program, pulses, times, 1 * u.asec, device=0)
    assert dev.shape == host.shape == (len(times) - 1, len(pulses))
    scale = np.max(np.abs(host), axis=0)
    assert np.max(np.abs(dev - host) / scale[None, :]) < 1e-12
    # anything that is not a plain windowed Sinc pulse goes pulse by pulse on the host (here: a DC-corrected pulse is a sum)
    corrected = P.DC_correct_electric_potential(pulses[0], times)
    assert C.sinc_pulse_table([corrected]) is None
