"""The reference-side binding of INTEGRATION.md as code: give it the imported reference package (``import ionization``) and it
routes the reference's own objects through the C-ABI.  Nothing here imports the reference; everything is duck-typed on the
objects it is handed, so the module loads (and is unit-tested) without it.

    import ionization, ionization_b200.reference_binding as rb
    rb.bind_tdma(ionization)                                   # level 1: cy.tdma -> ion_tdma_c128
    spec = ionization.mesh.SphericalHarmonicSpecification(..., evolution_method=rb.make_evolution_method(ionization)())
    spec.to_sim().run()                                        # level 2: every QuantumMesh.evolve() is one ion_sim_step

Level 3 (the device-resident run loop, where the speed is) is ``ionization_b200.mesh`` itself, which accepts the reference's
pulse / potential / mask / state objects unchanged.
Citations are relative to /root/reference/ionization/.
"""
import numpy as np

from . import engine as _engine
from . import exceptions


def bind_tdma(ionization):
    """cy.tdma (cy.pyx:9-50) -> ion_tdma_c128.  Its one production caller is TDMAOperator._apply (mesh/mesh_operators.py:109-111),
    which looks the function up as ``cy.tdma`` at call time, so rebinding the module attribute is the whole change."""
    ionization.cy.tdma = _engine.tdma
    return ionization.cy.tdma


def _dia_vectors(matrix, shape):
    """(diag [L, R], off [R-1]) of a scipy dia_matrix with offsets (-1, 0, 1) built R-wrapped (mesh_operators.py:889-928)"""
    L, R = shape
    offsets = list(matrix.offsets)
    diag = np.asarray(matrix.data[offsets.index(0)]).reshape(L, R)
    sup = np.asarray(matrix.data[offsets.index(1)])[1:]  # scipy dia: data[k][j] sits in column j
    off = np.concatenate([sup, [0]]).reshape(L, R)[0, :-1]
    return np.ascontiguousarray(diag, dtype=np.complex128), np.ascontiguousarray(np.real(off), dtype=np.float64)


def extract_problem(sim, units):
    """The hot-path inputs of a REFERENCE SphericalHarmonic simulation as the dictionary ``engine.DeviceSimulation.from_problem``
    takes (keys as tests/golden/*.npz), read out of the reference's own objects: H0 from operators.internal_hamiltonian(mesh)
    (mesh_operators.py:244-269), couplings per :988-1006 / :1143-1178, the mask evaluated once (potentials/masks.py:76-89)."""
    spec, mesh = sim.spec, sim.mesh
    ops = spec.operators
    L, R = mesh.mesh_shape
    (h0,) = ops.internal_hamiltonian(mesh).operators
    h_diag, h_off = _dia_vectors(h0.matrix, (L, R))
    l = np.arange(L - 1)
    c_l = np.asarray(ops.c_l(l), dtype=np.float64)
    q, m = spec.test_charge, spec.test_mass
    velocity = ops.__class__.__name__ == "SphericalHarmonicVelocityGaugeOperators"
    adi = spec.evolution_method.__class__.__name__.endswith("AlternatingDirectionImplicit")
    if velocity and adi:
        raise exceptions.UnsupportedConfiguration("velocity-gauge operators only work with SplitInteractionOperator (mesh_operators.py:1188)")
    p = dict(
        kind="sh_vel_so" if velocity else ("sh_len_adi" if adi else "sh_len_so"), L=L, R=R, r=np.asarray(mesh.r), delta_r=float(mesh.delta_r), h_diag=h_diag, h_off=h_off,
        c_l=c_l, mask=np.broadcast_to(np.asarray(spec.mask(r=mesh.r), dtype=np.float64), (R,)).copy(), g0=np.asarray(mesh.g, dtype=np.complex128),
        state_l=np.array([s.l for s in spec.test_states], dtype=np.int64), state_rows=np.array([mesh.get_radial_g_for_state(s) for s in spec.test_states]),
    )
    if velocity:
        p["f1_l"] = c_l * (l + 1)
        p["y_j"] = units.hbar * (q / m) / np.asarray(mesh.r)
        p["z_j"] = units.hbar * (q / m) / (2 * mesh.delta_r) * np.asarray(ops.alpha(np.arange(R - 1)), dtype=np.float64)
    else:
        p["x_j"] = -q * np.asarray(mesh.r)
    return p


def make_evolution_method(ionization, device=0):
    """-> a subclass of the reference's EvolutionMethod (mesh/evolution_methods.py:12-24) whose evolve(mesh, g, time_step) is one
    ion_sim_step on a handle built once per mesh from the reference's own objects.  evolve() excludes the mask: QuantumMesh.evolve
    applies it afterwards (meshes.py:256-257)."""
    import simulacra.units as u  # the reference's own unit module (it is imported whenever the reference is)

    base = ionization.mesh.evolution_methods.EvolutionMethod

    class B200SplitInteractionOperator(base):
        def __init__(self):
            self._handles = {}

        def _handle(self, mesh):
            key = id(mesh)
            if key not in self._handles:
                p = extract_problem(mesh.sim, u)
                h = _engine.DeviceSimulation.from_problem(p, device=device, with_states=False)
                h.set_mask(None)
                self._handles[key] = (h, p["kind"])
            return self._handles[key]

        def evolve(self, mesh, g, time_step):
            h, kind = self._handle(mesh)
            spec, sim = mesh.spec, mesh.sim
            if kind == "sh_vel_so":  # A over times[0 .. n+1] (mesh_operators.py:1184-1186)
                field = spec.electric_potential.get_vector_potential_amplitude_numeric(sim.times_to_current)
            else:  # E(t_{n+1} + dt/2) (mesh_operators.py:1011-1013)
                field = spec.electric_potential.get_electric_field_amplitude(sim.time + spec.time_step / 2)
            h.write_g(np.asarray(g, dtype=np.complex128).reshape(1, *mesh.mesh_shape))
            h.step(np.array([time_step / (2 * u.hbar)]), np.array([float(field)]))
            return h.read_g()[0].reshape(np.shape(g))

        def info(self):
            return ionization.mesh.evolution_methods.EvolutionMethod.info(self) if hasattr(base, "info") else None

        def __del__(self):
            for h, _ in getattr(self, "_handles", {}).values():
                try:
                    h.close()
                except Exception:  # noqa: BLE001
                    pass

    return B200SplitInteractionOperator
