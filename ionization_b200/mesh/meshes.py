"""Mesh state objects (ionization/mesh/meshes.py): coordinates, the wavefunction ``g`` and its observables.

The wavefunction lives on the GPU inside a ``DeviceSimulation``; ``mesh.g`` is a lazily synchronised, settable and
picklable host view (SURVEY.md 8b "ownership"): reading it after the device advanced copies it back once; assigning
it uploads it.  Observables of the CURRENT wavefunction (norm, inner products with the test states, <r>, <z>, <H0>,
norm by l, norm within a radius) are device reductions; the same quantities for an arbitrary host array passed by
the caller (e.g. when normalising a state at set-up) are plain numpy on that array.
"""
import numpy as np
import scipy.linalg

from .. import _native as nat
from .. import engine as _engine
from .. import exceptions, states
from .. import units as u
from ..core import WrappingDirection


class QuantumMesh:
    """meshes.py:55-260"""

    def __init__(self, sim):
        self.sim = sim
        self.spec = sim.spec
        self.operators = self.spec.operators
        self._g_host = None
        self._host_valid = True  # host copy is the newest
        self._device_valid = False  # device copy is the newest
        self._engine = None
        self.inner_product_multiplier = None

    # ---- wavefunction: lazily synchronised between host and device ---------------------------------
    @property
    def g(self):
        if not self._host_valid and self._engine is not None:
            self._g_host = self._engine.read_g()[0].reshape(self.mesh_shape)
            self._host_valid = True
        return self._g_host

    @g.setter
    def g(self, value):
        self._g_host = None if value is None else np.array(value, dtype=np.complex128).reshape(self.mesh_shape)
        self._host_valid = True
        self._device_valid = False

    def _mark_device_advanced(self):
        self._host_valid = False
        self._device_valid = True

    def _upload_if_needed(self):
        if not self._device_valid:
            self._engine.write_g(self._g_host.reshape(1, -1))
            self._device_valid = True

    @property
    def engine(self) -> "_engine.DeviceSimulation":
        if self._engine is None:
            self._engine = self._build_engine()
        return self._engine

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_g_host"] = self.g  # forces the device->host copy
        state["_host_valid"] = True
        state["_device_valid"] = False
        state["_engine"] = None
        return state

    def __eq__(self, other):
        return isinstance(other, self.__class__) and self.sim == other.sim and np.array_equal(self.g, other.g)

    def __hash__(self):
        return hash((self.__class__.__name__, self.sim))

    def __str__(self):
        return f"{self.__class__.__name__} for {self.sim}"

    # ---- flatten / wrap (meshes.py:117-133), kept for API compatibility --------------------------------
    def flatten_mesh(self, mesh, flatten_along):
        flat = self.wrapping_direction_to_order(flatten_along)
        return mesh if flat is None else mesh.flatten(flat)

    def wrap_vector(self, vector, wrap_along):
        wrap = self.wrapping_direction_to_order(wrap_along)
        return vector if wrap is None else np.reshape(vector, self.mesh_shape, wrap)

    def wrapping_direction_to_order(self, wrapping_direction):
        return None

    # ---- observables -------------------------------------------------------------------------------
    def state_to_g(self, state_or_mesh):
        if state_or_mesh is None:
            return self.g
        if isinstance(state_or_mesh, states.QuantumState):
            try:
                state_or_mesh = self.analytic_to_numeric[state_or_mesh]
            except (AttributeError, KeyError):
                pass
            return self.get_g_for_state(state_or_mesh)
        return state_or_mesh

    def get_g_with_states_removed(self, states_to_remove, g=None):
        """g - sum_s <s|g> |s>  (meshes.py:163-193); always acts on a copy"""
        g = np.array(self.state_to_g(g), dtype=np.complex128, copy=True)
        for state in states_to_remove:
            g -= self.inner_product(state, g) * self.get_g_for_state(state)
        return g

    def _observe(self, what):
        """device reductions of the current wavefunction -> dict"""
        eng = self.engine
        self._upload_if_needed()
        line_z = isinstance(self, LineMesh) and bool(what & nat.OBS_Z)
        if line_z:  # on a line <z> = sum z |g|^2 is the engine's "r" observable (there is no l coupling to evaluate)
            what = (what & ~nat.OBS_Z) | nat.OBS_R
        rec = self.sim._split_record(eng.observe(what)[0], what)
        if line_z:
            rec["z"] = rec["r"]
        return rec

    def inner_product(self, a=None, b=None):
        """meshes.py:195-200; (state, None) for a registered test state is a device reduction"""
        if b is None and isinstance(a, states.QuantumState) and a in self.sim._state_index:
            return self._observe(nat.OBS_INNER_PRODUCTS)["inner_products"][self.sim._state_index[a]]
        return np.sum(np.conj(self.state_to_g(a)) * self.state_to_g(b)) * self.inner_product_multiplier

    def state_overlap(self, a=None, b=None):
        return np.abs(self.inner_product(a, b)) ** 2

    def norm(self, state=None):
        """meshes.py:215-217"""
        if state is None:
            return float(self._observe(nat.OBS_NORM)["norm"])
        g = self.state_to_g(state)
        return float(np.real(np.sum(np.conj(g) * g) * self.inner_product_multiplier))

    def r_expectation_value(self, state=None):
        if state is None:
            return float(self._observe(nat.OBS_R)["r"])
        g = self.state_to_g(state)
        return float(np.real(np.sum(np.conj(g) * (self.r_mesh * g)) * self.inner_product_multiplier))

    def z_expectation_value(self, state=None):
        if state is not None:
            raise exceptions.UnsupportedConfiguration("z_expectation_value is only available for the current wavefunction")
        return float(self._observe(nat.OBS_Z)["z"])

    def internal_energy_expectation_value(self, state=None):
        if state is not None:
            raise exceptions.UnsupportedConfiguration("internal_energy_expectation_value is only available for the current wavefunction")
        return float(self._observe(nat.OBS_H0)["internal_energy"])

    def total_energy_expectation_value(self, state=None):
        if state is not None:
            raise exceptions.UnsupportedConfiguration("total_energy_expectation_value is only available for the current wavefunction")
        rec = self._observe(nat.OBS_H0 | nat.OBS_Z)
        return float(self.sim._total_energy(rec, self.sim.time_index))

    @property
    def psi(self):
        return self.g / self.g_factor

    @property
    def g2(self):
        return np.abs(self.g) ** 2

    @property
    def psi2(self):
        return np.abs(self.psi) ** 2

    # ---- evolution ------------------------------------------------------------------------------------
    def evolve(self, time_step):
        """QuantumMesh.evolve (meshes.py:251-257): evolution operators, then the mask.  One step on the device; the
        field scalar is the one the reference would sample at ``sim.time`` (already advanced by the caller)."""
        n = self.sim.time_index - 1
        self.sim._advance(n, n + 1, observe=False)

    def _evolve_operator_only(self, g, time_step):
        eng = self.engine
        self.g = g
        self._upload_if_needed()
        n = self.sim.time_index - 1
        eng.set_mask(None)
        try:
            eng.step(self.sim._taus[n : n + 1], self.sim._fields[n : n + 1])
        finally:
            eng.set_mask(self.sim._mask_vector)
        self._mark_device_advanced()
        return self.g


class LineMesh(QuantumMesh):
    """meshes.py:263-427"""

    mesh_storage_method = ("z",)

    def __init__(self, sim):
        super().__init__(sim)
        spec = self.spec
        self.z_mesh = np.linspace(-spec.z_bound, spec.z_bound, spec.z_points)
        self.delta_z = np.abs(self.z_mesh[1] - self.z_mesh[0])
        self.z_center_index = int(np.argmin(np.abs(self.z_mesh)))
        self.mesh_points = len(self.z_mesh)
        self.mesh_shape = (self.mesh_points,)
        self.inner_product_multiplier = self.delta_z
        self.g_factor = 1
        self._g_for_state_cache = {}
        if spec.use_numeric_eigenstates:
            self.analytic_to_numeric = self._get_numeric_eigenstate_basis(spec.number_of_numeric_eigenstates)
            spec.test_states = sorted(list(self.analytic_to_numeric.values()), key=lambda x: x.energy)
            spec.initial_state = self.analytic_to_numeric[spec.initial_state]
        self.g = self.get_g_for_state(spec.initial_state)

    z = property(lambda self: self.z_mesh)
    r = property(lambda self: self.z_mesh)
    r_mesh = property(lambda self: self.z_mesh)

    def get_g_for_state(self, state):
        if state in self._g_for_state_cache:
            return self._g_for_state_cache[state]
        if getattr(state, "analytic", False) and self.spec.use_numeric_eigenstates:
            state = getattr(self, "analytic_to_numeric", {}).get(state, state)
        g = np.asarray(state(self.z_mesh), dtype=np.complex128)
        g = g / np.sqrt(self.norm(g))
        g = g * state.amplitude
        self._g_for_state_cache[state] = g
        return g

    def _get_numeric_eigenstate_basis(self, number_of_eigenstates):
        """meshes.py:343-388 (ARPACK eigsh there; LAPACK tridiagonal eigensolver here -- same eigenpairs up to sign)"""
        hd, ho = self.operators.hamiltonian_vectors(self)
        if np.max(np.abs(np.imag(hd))) > 0:
            raise exceptions.UnsupportedConfiguration("numeric eigenstates need a real internal potential")
        k = int(number_of_eigenstates)
        vals, vecs = scipy.linalg.eigh_tridiagonal(np.real(hd[0]), ho, select="i", select_range=(0, k - 1))
        out = {}
        for nn, (val, vec) in enumerate(zip(vals, vecs.T)):
            vec = vec / np.sqrt(self.inner_product_multiplier * np.sum(np.abs(vec) ** 2))
            try:
                bound = states.Binding.BOUND
                analytic = self.spec.analytic_eigenstate_type.from_potential(
                    self.spec.internal_potential, self.spec.test_mass, n=nn + self.spec.analytic_eigenstate_type.smallest_n
                )
            except exceptions.IllegalQuantumState:
                bound = states.Binding.FREE
                analytic = states.OneDPlaneWave.from_energy(val, mass=self.spec.test_mass)
            out[analytic] = states.NumericOneDState(g=vec.astype(np.complex128), energy=val, binding=bound, corresponding_analytic_state=analytic)
        return out

    def _build_engine(self):
        sim, spec = self.sim, self.spec
        eng = _engine.DeviceSimulation(sim._program, 1, self.mesh_points, batch=1, device=sim.device)
        hd, ho = self.operators.hamiltonian_vectors(self)
        eng.set_hamiltonian(hd, ho)
        self.operators.configure_engine(self, eng)
        eng.set_mask(sim._mask_vector)
        rows = np.array([self.get_g_for_state(s) for s in spec.test_states]) if spec.test_states else None
        eng.set_observables(self.inner_product_multiplier, self.z_mesh, np.zeros(len(spec.test_states), dtype=np.int64), rows, sim._radii)
        return eng


class SphericalHarmonicMesh(QuantumMesh):
    """meshes.py:985-1525"""

    mesh_storage_method = ("l", "r")

    def __init__(self, sim):
        super().__init__(sim)
        spec = self.spec
        self.r = np.linspace(0, spec.r_bound, spec.r_points)
        self.delta_r = self.r[1] - self.r[0]
        self.r += self.delta_r / 2
        self.r_max = np.max(self.r)
        self.inner_product_multiplier = self.delta_r
        self.l = np.array(range(spec.l_bound), dtype=int)
        self.theta_points = self.phi_points = spec.theta_points
        self.mesh_points = len(self.r) * len(self.l)
        self.mesh_shape = (len(self.l), len(self.r))
        self._radial_cache = {}
        if spec.use_numeric_eigenstates:
            self.analytic_to_numeric = self.get_numeric_eigenstate_basis(spec.numeric_eigenstate_max_energy, spec.numeric_eigenstate_max_angular_momentum)
            spec.test_states = sorted(list(self.analytic_to_numeric.values()), key=lambda x: x.energy)
            if not spec.initial_state.numeric:
                spec.initial_state = self.analytic_to_numeric[spec.initial_state]
        self.g = self.get_g_for_state(spec.initial_state)

    @property
    def r_mesh(self):
        return np.broadcast_to(self.r[None, :], self.mesh_shape)

    @property
    def l_mesh(self):
        return np.broadcast_to(self.l[:, None], self.mesh_shape)

    @property
    def g_factor(self):
        return self.r

    def wrapping_direction_to_order(self, wrapping_direction):
        if wrapping_direction is None:
            return None
        if wrapping_direction == WrappingDirection.L:
            return "F"
        if wrapping_direction == WrappingDirection.R:
            return "C"
        raise exceptions.InvalidWrappingDirection(f"{wrapping_direction} is not a valid specifier for flatten_mesh (valid specifiers: 'l', 'r')")

    def get_g_for_state(self, state):
        if not (isinstance(state, states.QuantumState) and all(hasattr(s, "spherical_harmonic") for s in state)):
            raise NotImplementedError("States with non-definite angular momentum components are not currently supported by SphericalHarmonicMesh")
        g = np.zeros(self.mesh_shape, dtype=np.complex128)
        for s in state:
            if getattr(s, "analytic", False) and self.spec.use_numeric_eigenstates:
                s = getattr(self, "analytic_to_numeric", {}).get(s, s)
            g[s.l, :] += self.get_radial_g_for_state(s)
        return g

    def get_radial_g_for_state(self, state):
        """meshes.py:1090-1097"""
        key = (state, state.amplitude)
        if key not in self._radial_cache:
            g = np.asarray(state.radial_function(self.r) * self.g_factor, dtype=np.complex128)
            g = g / np.sqrt(self.norm(g))
            self._radial_cache[key] = g * state.amplitude
        return self._radial_cache[key]

    def inner_product(self, a=None, b=None):
        """meshes.py:1099-1131"""
        if b is None and isinstance(a, states.QuantumState) and a in self.sim._state_index:
            return self._observe(nat.OBS_INNER_PRODUCTS)["inner_products"][self.sim._state_index[a]]
        if b is None and isinstance(a, states.QuantumState) and all(hasattr(s, "spherical_harmonic") for s in a):
            g = self.g
            return sum(np.sum(np.conj(self.get_radial_g_for_state(s)) * g[s.l, :]) for s in a) * self.inner_product_multiplier
        return super().inner_product(a, b)

    # ---- spatial (r, theta) reconstruction and the analysis built on it: host-side numpy on the synchronised g, evaluated at
    # ---- snapshot / data times only (meshes.py:1456-1513; not part of the per-step path)
    @property
    def theta_calc(self):
        return np.linspace(0, u.pi, self.theta_points)

    @property
    def _sph_harm_l_theta_calc_mesh(self):
        """Y_l^0(theta) on (l, theta_calc): sqrt((2l+1)/(4 pi)) P_l(cos theta) (meshes.py:1485-1489 via scipy's sph_harm there)"""
        if getattr(self, "_ylt_cache", None) is None:
            import scipy.special as special

            l_mesh, theta_mesh = np.meshgrid(self.l, self.theta_calc, indexing="ij")
            self._ylt_cache = (np.sqrt((2 * l_mesh + 1) / (4 * u.pi)) * special.eval_legendre(l_mesh, np.cos(theta_mesh))).astype(np.complex128)
        return self._ylt_cache

    def reconstruct_spatial_mesh__calc(self, mesh):
        """(l, r) -> (r, theta)  (meshes.py:1498-1503)"""
        return np.einsum("lr,lt->rt", mesh, self._sph_harm_l_theta_calc_mesh)

    @property
    def space_g_calc(self):
        return self.reconstruct_spatial_mesh__calc(self.g)

    def get_radial_probability_current_density_mesh__spatial(self):
        """Im(conj(g) D_r g) on the (r, theta) mesh with the antisymmetric radial difference operator D_r
        (meshes.py:1358-1370; mesh_operators.py:1106-1127).  The reference's own call omits the mesh argument of
        r_probability_current__spatial (meshes.py:1359) and therefore raises TypeError there; this is its evident intent."""
        off = self.operators.r_probability_current_offdiagonal(self)  # [R - 1], couples r_j and r_{j+1} at fixed theta
        g_spatial = self.space_g_calc
        grad = np.zeros_like(g_spatial)
        grad[:-1, :] += off[:, None] * g_spatial[1:, :]
        grad[1:, :] -= off[:, None] * g_spatial[:-1, :]
        return np.imag(np.conj(g_spatial) * grad)

    def inner_product_with_plane_waves(self, thetas, wavenumbers, g=None):
        """<plane wave(theta, k) | g> for the Cartesian product of thetas and wavenumbers (meshes.py:1138-1190):
        sum_{l, r} sqrt(2/pi) r (-i^{l mod 4}) dr g[l, r] Y_l^0(theta) j_l(k r).  The reference evaluates the double loop over
        (theta, k) in Python; here the sum over r is done once per wavenumber and the sum over l is one matrix product."""
        import scipy.special as special

        if g is None:
            g = self.g
        g = np.asarray(g, dtype=np.complex128)
        l_mesh = self.l_mesh
        multiplier = np.sqrt(2 / u.pi) * self.g_factor * (-(1j ** (l_mesh % 4))) * self.inner_product_multiplier * g
        thetas, wavenumbers = np.array(thetas), np.array(wavenumbers)
        theta_mesh, wavenumber_mesh = np.meshgrid(thetas, wavenumbers, indexing="ij")
        lt, tt = np.meshgrid(self.l, thetas, indexing="ij")
        y_lt = np.sqrt((2 * lt + 1) / (4 * u.pi)) * special.eval_legendre(lt, np.cos(tt))  # Y_l^0(theta), [L, n_theta]
        b_lk = np.empty((len(self.l), len(wavenumbers)), dtype=np.complex128)
        for jj, k in enumerate(wavenumbers):
            b_lk[:, jj] = np.sum(multiplier * special.spherical_jn(l_mesh, np.real(k * self.r_mesh)), axis=1)
        return theta_mesh, wavenumber_mesh, y_lt.T @ b_lk

    def norm_by_l(self, state=None):
        """meshes.py:1133-1136"""
        if state is None:
            return np.array(self._observe(nat.OBS_NORM_BY_L)["norm_by_l"])
        g = self.state_to_g(state)
        return np.abs(np.sum(np.conj(g) * g, axis=1) * self.delta_r)

    def get_numeric_eigenstate_basis(self, max_energy, max_angular_momentum):
        """meshes.py:1281-1356: per-l eigenpairs of the discretised H0 with energy <= max_energy (ARPACK eigsh with a
        growing k there; LAPACK's tridiagonal eigensolver by value range here -- the same set, vectors up to sign)."""
        out = {}
        for l in range(max_angular_momentum + 1):
            hd, ho = self.operators.single_l_hamiltonian(self, l)
            if np.max(np.abs(np.imag(hd))) > 0:
                raise exceptions.UnsupportedConfiguration("numeric eigenstates need a real internal potential")
            hd = np.real(hd)
            lo = float(np.min(hd) - 2 * np.max(np.abs(ho)) - 1.0 * u.eV)
            vals, vecs = scipy.linalg.eigh_tridiagonal(hd, ho, select="v", select_range=(lo, max_energy))
            for val, vec in zip(vals, vecs.T):
                vec = vec / np.sqrt(self.inner_product_multiplier * np.sum(np.abs(vec) ** 2))
                vec = vec / self.g_factor
                if val > 0:
                    analytic = states.HydrogenCoulombState(energy=val, l=l)
                    binding = states.Binding.FREE
                else:
                    n_guess = int(np.sqrt(u.rydberg / np.abs(val)))
                    analytic = states.HydrogenBoundState(n=max(n_guess, 1) if max(n_guess, 1) > l else l + 1, l=l)
                    binding = states.Binding.BOUND
                out[analytic] = states.NumericSphericalHarmonicState(
                    g=vec.astype(np.complex128), l=l, m=0, energy=val, corresponding_analytic_state=analytic, binding=binding
                )
        return out

    def _build_sharded_engine(self, devices):
        """l-block shards on several GPUs of this process (mesh/sharded.py)"""
        from .. import coefficients as C
        from . import sharded

        sim, spec = self.sim, self.spec
        if sim._program not in ("sh_len_so", "sh_vel_so"):
            raise exceptions.UnsupportedConfiguration("devices=[...]: l-block sharding is available for the split-operator SphericalHarmonic programs")
        hd, ho = self.operators.hamiltonian_vectors(self)
        flat = sim._flat_states
        problem = dict(
            kind=sim._program, L=spec.l_bound, R=spec.r_points, r=self.r, delta_r=self.delta_r, h_diag=hd, h_off=ho, g0=self.g,
            mask=sim._mask_vector if sim._mask_vector is not None else np.ones(spec.r_points),
            state_l=np.array([s.l for s in flat], dtype=np.int64), state_rows=np.array([self.get_radial_g_for_state(s) for s in flat]) if flat else None,
        )
        if sim._program == "sh_vel_so":
            problem["c_l"], problem["f1_l"], problem["y_j"], problem["z_j"] = C.sh_vel_coupling(self.r, self.delta_r, spec.l_bound, spec.test_charge, spec.test_mass)
        else:
            problem["c_l"], problem["x_j"] = C.sh_len_coupling(self.r, spec.l_bound, spec.test_charge)
        eng = sharded.ShardedEngine(problem, devices, radii=sim._radii)
        if sim._mask_vector is None:
            eng.set_mask(None)
        return eng

    def _build_engine(self):
        sim, spec = self.sim, self.spec
        L, R = self.mesh_shape
        devices = getattr(spec, "devices", None)
        if devices is not None and len(devices) > 1:
            return self._build_sharded_engine(devices)
        eng = _engine.DeviceSimulation(sim._program, L, R, batch=1, device=sim.device)
        hd, ho = self.operators.hamiltonian_vectors(self)
        eng.set_hamiltonian(hd, ho)
        self.operators.configure_engine(self, eng)
        eng.set_mask(sim._mask_vector)
        flat_states = sim._flat_states
        rows = np.array([self.get_radial_g_for_state(s) for s in flat_states]) if flat_states else None
        eng.set_observables(self.inner_product_multiplier, self.r, np.array([s.l for s in flat_states], dtype=np.int64), rows, sim._radii)
        return eng
