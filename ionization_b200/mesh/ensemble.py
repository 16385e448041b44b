"""Scan ensembles: many simulations that differ only in their electric field, evolved as one batched device run.

The reference maps ``run(spec)`` over independent processes (``ionization_scans/scan_utils.py:638-663``,
``si.utils.multi_map``); a pulse-parameter scan is the cartesian product of pulse parameters over ONE mesh
(``ionization_scans/scan_mesh.py:40-68``).  Here the members share every coefficient vector on the device and
only their per-step field scalars differ, so a whole shard of the scan is a single ``DeviceSimulation`` with
``batch = len(specs)`` -- the batch dimension is just more CTAs for the same kernels.
"""
import copy
import uuid

import numpy as np

from .. import engine as _engine
from .. import exceptions
from ..core import Status
from . import data as data_mod
from . import meshes


def _compatible(a, b):
    keys = ("r_bound", "r_points", "l_bound", "z_bound", "z_points", "time_initial", "time_final", "test_mass", "test_charge", "store_data_every")
    for k in keys:
        if getattr(a, k, None) != getattr(b, k, None):
            return f"{k} differs"
    if callable(a.time_step) or callable(b.time_step) or a.time_step != b.time_step:
        return "time_step differs (or is callable)"
    if type(a.operators) is not type(b.operators) or type(a.evolution_method) is not type(b.evolution_method):
        return "operators / evolution method differ"
    if repr(a.mask) != repr(b.mask):
        return "mask differs"
    if a.initial_state != b.initial_state:
        return "initial_state differs"
    # the members share ONE Hamiltonian, one basis of test states and one observation record layout on the device
    if repr(a.internal_potential) != repr(b.internal_potential):
        return "internal_potential differs"
    for k in ("use_numeric_eigenstates", "numeric_eigenstate_max_energy", "numeric_eigenstate_max_angular_momentum", "number_of_numeric_eigenstates",
              "analytic_eigenstate_type"):
        if getattr(a, k, None) != getattr(b, k, None):
            return f"{k} differs"
    if a.datastore_types != b.datastore_types:
        return "datastore types differ"
    if [repr(s) for s in a.test_states] != [repr(s) for s in b.test_states] and not getattr(a, "use_numeric_eigenstates", False):
        return "test_states differ"
    return None


class MeshEnsemble:
    """``MeshEnsemble(specs).run()`` -> list of finished simulations (same objects a loop of ``spec.to_sim().run()``
    would give: ``sim.data.*`` filled at the data times, ``sim.mesh.g`` the final wavefunction)."""

    def __init__(self, specs, device=0):
        specs = list(specs)
        if not specs:
            raise exceptions.EngineError("empty ensemble")
        # compared BEFORE the first member's to_sim(), which may replace its states by the numeric basis
        for s in specs[1:]:
            why = _compatible(specs[0], s)
            if why:
                raise exceptions.UnsupportedConfiguration(f"ensemble members must share the mesh and time grid: {why}")
        for sp in specs:
            if sp.snapshot_times or sp.snapshot_indices or any(getattr(ds, "needs_mesh", False) for ds in sp.datastores):
                raise exceptions.UnsupportedConfiguration("snapshots and mesh-analysing datastores need the wavefunction on the host at intermediate times; "
                                                          "run such members one by one (spec.to_sim().run())")
        self.device = device
        specs[0].device = device
        first = specs[0].to_sim()
        self.sims = [first]
        for spec in specs[1:]:
            self.sims.append(self._clone_member(first, spec))
        self.batch = len(self.sims)
        # the per-step field scalars of all members: two kernels for the whole scan when the pulses are plain windowed Sinc
        # pulses (csrc/fields.cuh), else pulse by pulse on the host -- as each member's own to_sim() would
        from .. import coefficients as C

        pulses = [s.spec.electric_potential for s in self.sims]
        try:
            fields = C.field_series_batch(first._program, pulses, first.times, first.spec.time_step, device=device)
        except exceptions.NoCudaDevice:
            fields = C.field_series_batch(first._program, pulses, first.times, first.spec.time_step, device=None)
        for i, s in enumerate(self.sims):
            s._fields = np.ascontiguousarray(fields[:, i])

    @staticmethod
    def _clone_member(first, spec):
        from .. import coefficients as C
        from .. import potentials

        # every member gets the corrections MeshSimulation.__init__ applies to its own pulse (mesh/sims.py:53-75; scans
        # DC-correct by default, ionization_scans/scan_utils.py:517)
        if spec.electric_potential_dc_correction:
            spec.electric_potential = potentials.DC_correct_electric_potential(spec.electric_potential, first.times)
        if spec.electric_potential_fluence_correction:
            spec.electric_potential = potentials.FluenceCorrector(
                electric_potential=spec.electric_potential, times=first.times, target_fluence=list(spec.electric_potential)[0].fluence
            )
        sim = copy.copy(first)
        sim.uuid = uuid.uuid4()
        sim.name = spec.name
        sim.file_name = spec.file_name
        # the member keeps the (possibly numeric) states the first mesh produced
        spec.test_states = first.spec.test_states
        spec.initial_state = first.spec.initial_state
        spec.device = first.device
        sim.spec = spec
        sim._fields = None  # filled for all members at once (MeshEnsemble.__init__)
        sim._host_field_cache = None
        sim.mesh = copy.copy(first.mesh)
        sim.mesh.sim = sim
        sim.mesh.spec = spec
        sim.mesh._g_host = np.array(first.mesh._g_host, copy=True)
        sim.mesh._engine = None
        sim.data = data_mod.Data(sim)
        sim.datastores_by_type = {ds.__class__: copy.deepcopy(ds) for ds in spec.datastores}
        for ds in sim.datastores_by_type.values():
            ds.init(sim)
        sim.warnings = type(first.warnings)(list)
        sim.time_index = 0
        sim.data_time_index = 0
        return sim

    def run(self):
        first = self.sims[0]
        mesh = first.mesh
        B = self.batch
        # one engine with a batch dimension, configured like the first member's
        if isinstance(mesh, meshes.SphericalHarmonicMesh):
            L, R = mesh.mesh_shape
        else:
            L, R = 1, mesh.mesh_points
        eng = _engine.DeviceSimulation(first._program, L, R, batch=B, device=self.device)
        try:
            hd, ho = mesh.operators.hamiltonian_vectors(mesh)
            eng.set_hamiltonian(hd, ho)
            mesh.operators.configure_engine(mesh, eng)
            eng.set_mask(first._mask_vector)
            flat = first._flat_states
            if isinstance(mesh, meshes.SphericalHarmonicMesh):
                rows = np.array([mesh.get_radial_g_for_state(s) for s in flat]) if flat else None
                eng.set_observables(mesh.inner_product_multiplier, mesh.r, np.array([s.l for s in flat], dtype=np.int64), rows, first._radii)
            else:
                rows = np.array([mesh.get_g_for_state(s) for s in flat]) if flat else None
                eng.set_observables(mesh.inner_product_multiplier, mesh.z_mesh, np.zeros(len(flat), dtype=np.int64), rows, first._radii)
            g0 = np.stack([np.asarray(s.mesh._g_host).reshape(L, R) for s in self.sims])
            eng.write_g(g0)
            fields = np.ascontiguousarray(np.stack([s._fields for s in self.sims], axis=1))
            what = first._obs_mask()
            for s in self.sims:
                s.status = Status.RUNNING
            # data at time index 0
            rec0 = eng.observe(what)
            last = first.time_steps - 1
            obs = first.data_mask[1:].astype(np.uint8)
            recs = eng.run(first._taus, fields, obs, what) if last > 0 else np.zeros((0, B, rec0.shape[1]))
            g_final = eng.read_g()
            # a temporary single-member engine view so _split_record knows the record layout
            for b, s in enumerate(self.sims):
                s.mesh._engine = eng
                s.time_index = 0
                s.data_time_index = 0
                s.store_data(s._split_record(rec0[b], what))
                s.check()
                s.data_time_index = 1
                k = 0
                for n in range(1, first.time_steps):
                    if first.data_mask[n]:
                        s.time_index = n
                        s.store_data(s._split_record(recs[k, b], what))
                        s.check()
                        s.data_time_index += 1
                        k += 1
                s.time_index = last
                s.mesh._engine = None
                s.mesh.g = g_final[b].reshape(s.mesh.mesh_shape)
                s.status = Status.FINISHED
        finally:
            eng.close()
        return self.sims


def run_ensemble(specs, device=0, devices=None):
    """``devices``: split the members into contiguous blocks, one batched device run per GPU, concurrently
    (ionization_b200.scan.run_scan; under torchrun the ranks take the blocks instead)"""
    if devices is not None and len(devices) > 1:
        from .. import scan

        return scan.run_scan(specs, devices=devices, keep_mesh=True)
    return MeshEnsemble(specs, device=devices[0] if devices else device).run()
