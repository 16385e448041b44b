"""Specifications and simulations (ionization/mesh/sims.py) with a device-resident time loop.

``MeshSimulation.run()`` keeps the order of the reference's loop (mesh/sims.py:289-333: store data -> check ->
callback -> advance) but, when nothing needs the host between data times (no callback, no animators, no
checkpoint due), it hands whole stretches of time steps to the CUDA engine in one call: the per-step field
scalars are precomputed exactly as the reference samples them, the wavefunction stays in HBM,
and the datastore values come back as one block of fused device reductions.
"""
import collections
import datetime
import functools
import logging
import operator
import os
import pickle
import uuid
from copy import deepcopy

import numpy as np

from .. import _native as nat
from .. import coefficients as C
from .. import exceptions, potentials, states
from .. import units as u
from ..core import Status
from . import data as data_mod
from . import evolution_methods, meshes, snapshots
from . import operators as mesh_operators

logger = logging.getLogger(__name__)

WarningRecord = collections.namedtuple("WarningRecord", ["time_index", "message"])


# ---------------------------------------------------------------------------------------------
# minimal Specification / Simulation bases (the reference inherits them from simulacra)
# ---------------------------------------------------------------------------------------------
class _Beet:
    def __init__(self, name, file_name=None):
        self.name = str(name)
        self.file_name = file_name or self.name
        self.uuid = uuid.uuid4()

    def __eq__(self, other):
        return isinstance(other, self.__class__) and self.uuid == other.uuid

    def __hash__(self):
        return hash(self.uuid)

    def __str__(self):
        return f"{self.__class__.__name__}({self.name})"

    __repr__ = __str__

    def _save(self, target_dir, extension):
        path = os.path.join(str(target_dir or os.getcwd()), f"{self.file_name}.{extension}")
        tmp = path + ".working"
        with open(tmp, "wb") as f:
            pickle.dump(self, f, protocol=-1)
        os.replace(tmp, path)  # atomic, as simulacra does
        return path

    @classmethod
    def load(cls, path):
        with open(str(path), "rb") as f:
            return pickle.load(f)


class MeshSpecification(_Beet):
    """mesh/sims.py:441-583"""

    simulation_type = None  # set below
    mesh_type = meshes.QuantumMesh

    def __init__(
        self,
        name,
        test_mass=u.electron_mass_reduced,
        test_charge=u.electron_charge,
        initial_state=None,
        test_states=tuple(),
        internal_potential=None,
        electric_potential=None,
        electric_potential_dc_correction=False,
        electric_potential_fluence_correction=False,
        mask=None,
        operators=None,
        evolution_method=None,
        time_initial=0 * u.asec,
        time_final=200 * u.asec,
        time_step=1 * u.asec,
        checkpoints=False,
        checkpoint_every=datetime.timedelta(hours=1),
        checkpoint_dir=None,
        animators=tuple(),
        store_data_every=1,
        snapshot_times=(),
        snapshot_indices=(),
        snapshot_type=None,
        snapshot_kwargs=None,
        datastores=None,
        device=0,
        devices=None,
        file_name=None,
        **kwargs,
    ):
        super().__init__(name, file_name=file_name)
        for k, v in kwargs.items():  # unknown kwargs become attributes (simulacra.Specification)
            setattr(self, k, v)
        self.test_mass = test_mass
        self.test_charge = test_charge
        self.initial_state = initial_state if initial_state is not None else states.HydrogenBoundState(1, 0)
        self.test_states = sorted(test_states)
        if len(self.test_states) == 0:
            self.test_states = [self.initial_state]
        self.internal_potential = internal_potential if internal_potential is not None else potentials.CoulombPotential(charge=u.proton_charge)
        self.electric_potential = electric_potential if electric_potential is not None else potentials.NoElectricPotential()
        self.electric_potential_dc_correction = electric_potential_dc_correction
        self.electric_potential_fluence_correction = electric_potential_fluence_correction
        self.mask = mask if mask is not None else potentials.NoMask()
        self.operators = operators
        self.evolution_method = evolution_method
        self.time_initial = time_initial
        self.time_final = time_final
        self.time_step = time_step
        self.checkpoints = checkpoints
        self.checkpoint_every = checkpoint_every
        self.checkpoint_dir = checkpoint_dir
        if len(tuple(animators)) > 0:
            raise exceptions.UnsupportedConfiguration("animators are visualisation (out of scope, SURVEY.md section 2); use run(callback=...)")
        self.animators = ()
        self.store_data_every = int(store_data_every)
        self.snapshot_times = set(snapshot_times)
        self.snapshot_indices = set(snapshot_indices)
        self.snapshot_type = snapshot_type if snapshot_type is not None else snapshots.Snapshot  # sims.py:565-567
        self.snapshot_kwargs = snapshot_kwargs or dict()
        if datastores is None:
            datastores = [ds_type() for ds_type in data_mod.DEFAULT_DATASTORE_TYPES]
        self.datastores = list(datastores)
        self.datastore_types = tuple(sorted(set(ds.__class__ for ds in self.datastores), key=lambda ds: ds.__name__))
        if len(self.datastores) != len(self.datastore_types):
            raise exceptions.DuplicateDatastores("Cannot duplicate datastores")
        # devices=[d0, d1, ...]: ONE simulation l-block sharded over several GPUs of this process (mesh/sharded.py); device: the GPU
        # of an unsharded simulation
        self.devices = None if devices is None else [int(d) for d in devices]
        self.device = int(device) if self.devices is None else self.devices[0]

    def to_sim(self):
        return self.simulation_type(self)

    def save(self, target_dir=None):
        return self._save(target_dir, "spec")


class LineSpecification(MeshSpecification):
    """mesh/sims.py:683-735"""

    mesh_type = meshes.LineMesh

    def __init__(
        self,
        name,
        internal_potential=None,
        initial_state=None,
        z_bound=10 * u.nm,
        z_points=2 ** 9,
        use_numeric_eigenstates=False,
        number_of_numeric_eigenstates=100,
        analytic_eigenstate_type=None,
        operators=None,
        evolution_method=None,
        **kwargs,
    ):
        super().__init__(
            name,
            internal_potential=internal_potential if internal_potential is not None else potentials.HarmonicOscillator(1 * u.N / u.m),
            initial_state=initial_state if initial_state is not None else states.QHOState(1 * u.N / u.m),
            operators=operators if operators is not None else mesh_operators.LineLengthGaugeOperators(),
            evolution_method=evolution_method if evolution_method is not None else evolution_methods.AlternatingDirectionImplicit(),
            **kwargs,
        )
        self.z_bound = z_bound
        self.z_points = int(z_points)
        self.analytic_eigenstate_type = analytic_eigenstate_type
        self.use_numeric_eigenstates = use_numeric_eigenstates
        self.number_of_numeric_eigenstates = number_of_numeric_eigenstates


class SphericalHarmonicSpecification(MeshSpecification):
    """mesh/sims.py:993-1059"""

    mesh_type = meshes.SphericalHarmonicMesh

    def __init__(
        self,
        name,
        r_bound=100 * u.bohr_radius,
        r_points=1000,
        l_bound=300,
        theta_points=180,
        operators=None,
        evolution_method=None,
        use_numeric_eigenstates=True,
        numeric_eigenstate_max_energy=20 * u.eV,
        numeric_eigenstate_max_angular_momentum=5,
        **kwargs,
    ):
        super().__init__(
            name,
            operators=operators if operators is not None else mesh_operators.SphericalHarmonicLengthGaugeOperators(),
            evolution_method=evolution_method if evolution_method is not None else evolution_methods.SplitInteractionOperator(),
            **kwargs,
        )
        self.r_bound = r_bound
        self.r_points = int(r_points)
        self.l_bound = int(l_bound)
        self.theta_points = theta_points
        self.spherical_harmonics = tuple(states.SphericalHarmonic(l, 0) for l in range(self.l_bound))
        self.use_numeric_eigenstates = use_numeric_eigenstates
        self.numeric_eigenstate_max_angular_momentum = min(self.l_bound - 1, numeric_eigenstate_max_angular_momentum)
        self.numeric_eigenstate_max_energy = numeric_eigenstate_max_energy


# ---------------------------------------------------------------------------------------------
# simulations
# ---------------------------------------------------------------------------------------------
class MeshSimulation(_Beet):
    """mesh/sims.py:36-438"""

    def __init__(self, spec):
        super().__init__(spec.name, file_name=spec.file_name)
        self.spec = spec
        self.status = Status.INITIALIZED
        self.device = getattr(spec, "device", 0)
        self.latest_checkpoint_time = datetime.datetime.now(datetime.timezone.utc)

        self.times = self.get_times()
        if spec.electric_potential_dc_correction:
            spec.electric_potential = potentials.DC_correct_electric_potential(spec.electric_potential, self.times)
        if spec.electric_potential_fluence_correction:
            spec.electric_potential = potentials.FluenceCorrector(
                electric_potential=spec.electric_potential, times=self.times, target_fluence=list(spec.electric_potential)[0].fluence
            )

        self.time_index = 0
        self.data_time_index = 0
        self.time_steps = len(self.times)

        self._program = spec.operators.program(spec.evolution_method)
        self._radii = ()
        for ds in spec.datastores:
            if isinstance(ds, data_mod.NormWithinRadius):
                self._radii = tuple(ds.radii)

        self.mesh = spec.mesh_type(self)

        time_indices = np.array(range(0, self.time_steps))
        self.data_mask = np.equal(time_indices, 0) + np.equal(time_indices, self.time_steps - 1)
        if spec.store_data_every >= 1:
            self.data_mask += np.equal(time_indices % spec.store_data_every, 0)
        self.data_times = self.times[self.data_mask]
        self.data_indices = time_indices[self.data_mask]
        self.data_time_steps = len(self.data_times)
        self.spacetime_points = self.time_steps * functools.reduce(operator.mul, self.mesh.mesh_shape)

        # test states: single-l components are the rows the engine projects on
        self._flat_states = []
        self._state_components = {}
        for s in spec.test_states:
            comps = []
            for c in s:
                comps.append(len(self._flat_states))
                self._flat_states.append(c)
            self._state_components[s] = comps
        self._state_index = {s: comps[0] for s, comps in self._state_components.items() if len(comps) == 1}

        # per-step scalars, exactly as the reference samples them
        self._taus = C.taus_from_times(self.times)
        self._fields = C.field_series(self._program, spec.electric_potential, self.times, spec.time_step)
        self._mask_vector = self._evaluate_mask()

        self.data = data_mod.Data(self)
        self.datastores_by_type = {ds.__class__: deepcopy(ds) for ds in spec.datastores}
        for ds in self.datastores_by_type.values():
            ds.init(self)
        self._what = 0
        for ds in self.datastores_by_type.values():
            self._what |= ds.observables
        if isinstance(self, SphericalHarmonicSimulation):
            self._what |= nat.OBS_NORM_BY_L | nat.OBS_NORM
        self._what |= nat.OBS_NORM  # check() needs it
        self._host_field_cache = None

        # snapshot times from the two ways of entering them in the spec, by time or by index (sims.py:104-114)
        self.snapshot_times = set()
        self._snapshot_indices = set()
        for t in spec.snapshot_times:
            idx = int(np.argmin(np.abs(self.times - t)))  # simulacra.utils.find_nearest_entry
            self.snapshot_times.add(self.times[idx])
            self._snapshot_indices.add(idx)
        for idx in spec.snapshot_indices:
            self.snapshot_times.add(self.times[idx])
            self._snapshot_indices.add(int(idx) % self.time_steps)
        self.snapshots = dict()
        self._needs_mesh = any(ds.needs_mesh for ds in self.datastores_by_type.values())
        self.warnings = collections.defaultdict(list)

    # ---- helpers -----------------------------------------------------------------------------------
    def _evaluate_mask(self):
        m = self.spec.mask(r=self.mesh.r)
        if np.ndim(m) == 0:
            return None if m == 1 else np.full(len(self.mesh.r), float(m))
        return np.asarray(m, dtype=np.float64)

    def get_blank_data(self, dtype=np.float64):
        a = np.empty(self.data_time_steps, dtype=dtype)
        a.fill(np.nan)
        return a

    @property
    def time(self):
        return self.times[self.time_index]

    @property
    def times_to_current(self):
        return self.times[: self.time_index + 1]

    def get_times(self):
        return C.time_grid(self.spec.time_initial, self.spec.time_final, self.spec.time_step, self.spec)

    @property
    def percent_completed(self):
        return round(100 * self.time_index / (self.time_steps - 1), 2)

    @property
    def bound_states(self):
        """sims.py:362-364"""
        yield from (s for s in self.spec.test_states if s.bound)

    @property
    def free_states(self):
        """sims.py:366-368"""
        yield from (s for s in self.spec.test_states if not s.bound)

    def take_snapshot(self):
        """sims.py:242-253"""
        snapshot = self.spec.snapshot_type(self, self.time_index, **self.spec.snapshot_kwargs)
        snapshot.take_snapshot()
        self.snapshots[self.time_index] = snapshot
        logger.info(f"Stored {snapshot.__class__.__name__} for {self} at time index {self.time_index} (t = {self.time / u.asec:.3f} as)")

    def _split_record(self, rec, what):
        """engine record (include/ionization_b200.h: ion_sim_observation_size) -> dict"""
        eng = self.mesh.engine
        out = {}
        c = 0
        if what & nat.OBS_NORM:
            out["norm"] = rec[c]
            c += 1
        if what & nat.OBS_INNER_PRODUCTS:
            n = eng.n_states
            flat = rec[c : c + 2 * n].reshape(n, 2)
            flat = flat[:, 0] + 1j * flat[:, 1]
            c += 2 * n
            out["inner_products_flat"] = flat
            out["inner_products"] = np.array([sum(flat[i] for i in self._state_components[s]) for s in self.spec.test_states])
        if what & nat.OBS_NORM_BY_L:
            out["norm_by_l"] = rec[c : c + eng.L]
            c += eng.L
        if what & nat.OBS_R:
            out["r"] = rec[c]
            c += 1
        if what & nat.OBS_Z:
            out["z"] = rec[c]
            c += 1
        if what & nat.OBS_H0:
            out["internal_energy"] = rec[c]
            c += 1
        if what & nat.OBS_NORM_WITHIN:
            out["norm_within_radius"] = rec[c : c + eng.n_radii]
            c += eng.n_radii
        return out

    def _total_energy(self, record, time_index):
        """<H0> + <H_int> at time index ``time_index`` (meshes.py:225-229).  Length gauge only: the reference cannot
        evaluate it in the velocity gauge either (SumOfOperators of raw matrices, mesh_operators.py:1188)."""
        spec = self.spec
        t = self.times[time_index]
        if self._program in ("sh_len_so", "sh_len_adi"):
            e = spec.electric_potential.get_electric_field_amplitude(t + spec.time_step / 2)  # mesh_operators.py:1011-1013
            return record["internal_energy"] + e * (-spec.test_charge) * record["z"]
        if self._program in ("line_len_cn", "line_len_so"):
            e = spec.electric_potential.get_electric_field_amplitude(t)  # :321-323 ; <z> is the "r" observable on a line
            return record["internal_energy"] + e * (-spec.test_charge) * record["z"]
        raise exceptions.UnsupportedConfiguration("total energy is not available in the velocity gauge (nor in the reference)")

    def _host_fields(self):
        """E(t) and A(t) at every time index for the Fields datastore (data.py:160-170)"""
        if self._host_field_cache is None:
            pot = self.spec.electric_potential
            e = np.asarray(pot.get_electric_field_amplitude(self.times), dtype=np.float64) * np.ones(self.time_steps)
            a = np.concatenate([[-potentials.simps(e[:1], self.times[:1])], C.vector_potential_series(pot, self.times)]) if self.time_steps > 1 else np.zeros(1)
            self._host_field_cache = (e, a)
        return self._host_field_cache

    def _complete_record(self, record, time_index):
        if data_mod.Fields in self.datastores_by_type:
            e, a = self._host_fields()
            record["electric_field_amplitude"] = e[time_index]
            record["vector_potential_amplitude"] = a[time_index]
        if self.mesh.__class__ is meshes.LineMesh and "z" not in record and "r" in record:
            record["z"] = record["r"]  # on a line <z> is the engine's "r" observable (see _obs_mask); total energy needs it
        if data_mod.TotalEnergyExpectationValue in self.datastores_by_type:
            record["total_energy"] = self._total_energy(record, time_index)
        return record

    def store_data(self, record=None):
        """mesh/sims.py:222-226"""
        if record is None:
            record = self.mesh._observe(self._obs_mask())
        record = self._complete_record(record, self.time_index)
        for ds in self.datastores_by_type.values():
            ds.store(record, self.data_time_index)
        self._last_record = record

    def _obs_mask(self):
        what = self._what
        if self.mesh.__class__ is meshes.LineMesh:
            # on a line <z> is sum z |g|^2, the engine's "r" observable
            if what & nat.OBS_Z:
                what = (what & ~nat.OBS_Z) | nat.OBS_R
            what &= ~nat.OBS_NORM_BY_L
        return what

    def check(self):
        """mesh/sims.py:228-240"""
        norm = self.data.norm[self.data_time_index] if data_mod.Norm in self.datastores_by_type else self._last_record["norm"]
        norm0 = self.data.norm[0] if data_mod.Norm in self.datastores_by_type else norm
        if norm > 1.001 * norm0:
            logger.warning(f"Wavefunction norm ({norm}) has exceeded initial norm ({norm0}) by more than .1% for {self}")

    # ---- evolution ---------------------------------------------------------------------------------
    def _advance(self, n0, n1, observe):
        """advance steps n0..n1-1 on the device.  observe: None/False, or uint8 mask per step -> records"""
        mesh = self.mesh
        eng = mesh.engine
        mesh._upload_if_needed()
        if observe is None or observe is False:
            eng.step(self._taus[n0:n1], self._fields[n0:n1])
            mesh._mark_device_advanced()
            return None
        what = self._obs_mask()
        recs = eng.run(self._taus[n0:n1], self._fields[n0:n1], observe, what)
        mesh._mark_device_advanced()
        return [self._split_record(r[0], what) for r in recs]

    def run(self, progress_bar=False, callback=None, checkpoint_callback=None):
        """MeshSimulation.run (mesh/sims.py:255-346)"""
        if checkpoint_callback is None:
            checkpoint_callback = lambda p: None
        self.status = Status.RUNNING
        last = self.time_steps - 1
        interactive = callback is not None
        # maximum number of steps handed to the device in one call when checkpoints may be due
        chunk_limit = 500 if self.spec.checkpoints else None

        while True:
            if self.data_mask[self.time_index]:  # same truth values as `self.time in self.data_times` (:290), O(1)
                self.store_data()
                self.check()
            if self.time_index in self._snapshot_indices:  # `self.time in self.snapshot_times` (sims.py:296-297)
                self.take_snapshot()
            if callback is not None:
                callback(self)
            if self.data_mask[self.time_index]:
                self.data_time_index += 1
            if self.time_index == last:
                break

            if interactive:
                # a callback may read sim.mesh.g / sim.data after every step: one step at a time
                self.time_index += 1
                self.mesh.evolve(self.times[self.time_index] - self.times[self.time_index - 1])
            else:
                # device-resident stretch up to the end (or the next checkpoint opportunity)
                n0 = self.time_index
                n1 = last if chunk_limit is None else min(last, n0 + chunk_limit)
                # the host needs the wavefunction at snapshot times, and at every data time when a datastore analyses the mesh itself:
                # the stretch ends there and the loop head above does the work
                stops = [i for i in self._snapshot_indices if i > n0]
                if self._needs_mesh:
                    stops += [int(i) for i in self.data_indices if i > n0]
                if stops:
                    n1 = min(n1, min(stops))
                obs = self.data_mask[n0 + 1 : n1 + 1].astype(np.uint8)
                obs[-1] = 0  # the final index of the stretch is stored by the loop head above
                recs = self._advance(n0, n1, obs)
                k = 0
                for n in range(n0 + 1, n1):
                    if self.data_mask[n]:
                        self.time_index = n
                        self.store_data(recs[k])
                        self.check()
                        self.data_time_index += 1
                        k += 1
                self.time_index = n1

            if self.spec.checkpoints:
                now = datetime.datetime.now(datetime.timezone.utc)
                if (now - self.latest_checkpoint_time) > self.spec.checkpoint_every:
                    self.do_checkpoint(now, checkpoint_callback)

        self.status = Status.FINISHED
        return self

    def do_checkpoint(self, now, callback):
        """mesh/sims.py:348-356"""
        self.status = Status.PAUSED
        path = self.save(target_dir=self.spec.checkpoint_dir, save_mesh=True)
        callback(path)
        self.latest_checkpoint_time = now
        self.status = Status.RUNNING

    def save(self, target_dir=None, save_mesh=True):
        """mesh/sims.py:405-438: atomically pickle to {target_dir}/{name}.sim"""
        mesh = self.mesh
        if not save_mesh:
            for state in self.spec.test_states:
                if hasattr(state, "g"):
                    state.g = None
            self.mesh = None
        try:
            return self._save(target_dir, "sim")
        finally:
            self.mesh = mesh

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_last_record", None)
        return state


class SphericalHarmonicSimulation(MeshSimulation):
    """mesh/sims.py:973-990"""

    def check(self):
        super().check()
        rec = getattr(self, "_last_record", None)
        if rec is None or "norm_by_l" not in rec:
            return
        norm_in_largest_l = abs(rec["norm_by_l"][-1]) ** 2  # state_overlap(g[-1], g[-1]) squares it
        if norm_in_largest_l > 1e-6:
            msg = (
                f"Wavefunction norm in largest angular momentum state is large at time index {self.time_index} "
                f"(norm at bound = {norm_in_largest_l}), consider increasing l bound"
            )
            logger.warning(msg)
            self.warnings["norm_in_largest_l"].append(WarningRecord(self.time_index, msg))


MeshSpecification.simulation_type = MeshSimulation
LineSpecification.simulation_type = MeshSimulation
SphericalHarmonicSpecification.simulation_type = SphericalHarmonicSimulation
