"""Operator strategy objects: same class names and constructor arguments as the reference
(ionization/mesh/mesh_operators.py), but instead of materialising scipy sparse matrices every time step they
produce the coefficient vectors and per-step scalars consumed by the CUDA engine (ionization_b200.coefficients).
"""
import numpy as np

from .. import coefficients as C
from .. import exceptions
from ..core import Gauge, KineticEnergyDerivation


class MeshOperators:
    gauge = None
    mesh_kind = None  # "sh" | "line"

    def __repr__(self):
        return f"{self.__class__.__name__}()"

    def info(self):
        return self.__class__.__name__

    def program(self, evolution_method) -> str:
        """engine program name for (these operators, evolution_method)"""
        key = (self.mesh_kind, self.gauge, evolution_method.kind)
        try:
            return {
                ("sh", Gauge.LENGTH, "so"): "sh_len_so",
                ("sh", Gauge.VELOCITY, "so"): "sh_vel_so",
                ("sh", Gauge.LENGTH, "adi"): "sh_len_adi",
                ("line", Gauge.LENGTH, "adi"): "line_len_cn",
                ("line", Gauge.LENGTH, "so"): "line_len_so",
                ("line", Gauge.VELOCITY, "so"): "line_vel_so",
            }[key]
        except KeyError:
            raise exceptions.UnsupportedConfiguration(
                f"{self.__class__.__name__} with {evolution_method.__class__.__name__} is not available "
                "(the reference cannot run it either: velocity-gauge operators only work with SplitInteractionOperator, "
                "mesh_operators.py:1188)"
            )


class SphericalHarmonicLengthGaugeOperators(MeshOperators):
    """mesh_operators.py:815-1127"""

    gauge = Gauge.LENGTH
    mesh_kind = "sh"

    def __init__(self, kinetic_energy_derivation=KineticEnergyDerivation.LAGRANGIAN, hydrogen_zero_angular_momentum_correction: bool = True):
        if KineticEnergyDerivation(kinetic_energy_derivation) != KineticEnergyDerivation.LAGRANGIAN:
            # the HAMILTONIAN derivation couples adjacent l blocks in the reference: not reproduced
            raise exceptions.UnsupportedConfiguration("only KineticEnergyDerivation.LAGRANGIAN is supported")
        self.kinetic_energy_derivation = KineticEnergyDerivation.LAGRANGIAN
        self.hydrogen_zero_angular_momentum_correction = hydrogen_zero_angular_momentum_correction

    alpha = staticmethod(C.sh_alpha)
    beta = staticmethod(C.sh_beta)
    c_l = staticmethod(C.sh_c_l)

    @staticmethod
    def gamma(j):
        """for the radial probability current (mesh_operators.py:849-851)"""
        return 1 / ((np.asarray(j, dtype=np.float64) ** 2) - 0.25)

    def r_probability_current_offdiagonal(self, mesh):
        """the super-diagonal of the antisymmetric radial-current operator along r, [r_points - 1]: hbar / (2 m dr^3) gamma(j),
        j = 1 .. r_points - 1; the sub-diagonal is its negative (mesh_operators.py:1106-1127)"""
        from .. import units as u

        pre = u.hbar / (2 * mesh.spec.test_mass * (mesh.delta_r ** 3))
        return pre * self.gamma(np.arange(1, mesh.spec.r_points))

    def hamiltonian_vectors(self, mesh):
        spec = mesh.spec
        V = spec.internal_potential(r=mesh.r, test_charge=spec.test_charge)
        return C.sh_hamiltonian(mesh.r, mesh.delta_r, spec.l_bound, V, self.hydrogen_zero_angular_momentum_correction)

    def single_l_hamiltonian(self, mesh, l):
        """(diag, off) of H0 for one channel (internal_hamiltonian_for_single_l, :959-978)"""
        spec = mesh.spec
        V = spec.internal_potential(r=mesh.r, test_charge=spec.test_charge)
        hd, ho = C.sh_hamiltonian(mesh.r, mesh.delta_r, 1, V, self.hydrogen_zero_angular_momentum_correction, l_begin=l)
        return hd[0], ho

    def configure_engine(self, mesh, engine_sim):
        spec = mesh.spec
        engine_sim.set_len_coupling(*C.sh_len_coupling(mesh.r, spec.l_bound, spec.test_charge))


class SphericalHarmonicVelocityGaugeOperators(SphericalHarmonicLengthGaugeOperators):
    """mesh_operators.py:1130-1408"""

    gauge = Gauge.VELOCITY

    def __init__(self, hydrogen_zero_angular_momentum_correction: bool = True):
        super().__init__(hydrogen_zero_angular_momentum_correction=hydrogen_zero_angular_momentum_correction)

    def configure_engine(self, mesh, engine_sim):
        spec = mesh.spec
        engine_sim.set_vel_coupling(*C.sh_vel_coupling(mesh.r, mesh.delta_r, spec.l_bound, spec.test_charge, spec.test_mass))


class LineLengthGaugeOperators(MeshOperators):
    """mesh_operators.py:304-349"""

    gauge = Gauge.LENGTH
    mesh_kind = "line"

    def hamiltonian_vectors(self, mesh):
        spec = mesh.spec
        V = spec.internal_potential(r=mesh.z_mesh, z=mesh.z_mesh, test_charge=spec.test_charge)
        hd, ho = C.line_hamiltonian(mesh.z_mesh, mesh.delta_z, V, spec.test_mass)
        return hd.reshape(1, -1), ho

    def configure_engine(self, mesh, engine_sim):
        spec = mesh.spec
        w_z, v_pref = C.line_coupling(mesh.z_mesh, mesh.delta_z, spec.test_charge, spec.test_mass)
        engine_sim.set_line_coupling(w_z, v_pref)


class LineVelocityGaugeOperators(LineLengthGaugeOperators):
    """mesh_operators.py:352-427 (the reference mislabels its gauge attribute as LENGTH, :355)"""

    gauge = Gauge.VELOCITY
