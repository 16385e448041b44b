"""Quantum states used to initialise the wavefunction and as projection targets (host-side inputs).

Mirrors the parts of ionization/states/ that the mesh path touches: ordering / hashing by ``tuple``
(states/state.py:150-175), amplitudes, bound/free flags, ``radial_function`` for hydrogen
(states/three_d.py:375-391), the 1-D states of the LineMesh configs (states/one_d.py) and the numeric
eigenstates produced by the meshes (states/three_d.py NumericSphericalHarmonicState, one_d.py NumericOneDState).
"""
import collections
from copy import deepcopy

import numpy as np
import scipy.optimize as optimize
import scipy.special as special

from . import exceptions
from . import units as u
from .core import StrEnum
from .potentials import Sum, Summand


class Eigenvalues(StrEnum):
    DISCRETE = "discrete"
    CONTINUOUS = "continuous"


class Binding(StrEnum):
    BOUND = "bound"
    FREE = "free"


class Derivation(StrEnum):
    ANALYTIC = "analytic"
    NUMERIC = "numeric"
    VARIATIONAL = "variational"


class SphericalHarmonic:
    """stand-in for simulacra.math.SphericalHarmonic (hashable, orderable, callable)"""

    def __init__(self, l=0, m=0):
        self.l = l
        self.m = m

    def __hash__(self):
        return hash((self.l, self.m))

    def __eq__(self, other):
        return isinstance(other, SphericalHarmonic) and (self.l, self.m) == (other.l, other.m)

    def __lt__(self, other):
        return (self.l, self.m) < (other.l, other.m)

    def __repr__(self):
        return f"SphericalHarmonic(l={self.l}, m={self.m})"

    def __call__(self, theta, phi=0):
        return special.sph_harm_y(self.l, self.m, theta, phi)


class QuantumState(Summand):
    """states/state.py:51-215"""

    eigenvalues = None
    binding = None
    derivation = None

    def __init__(self, amplitude=1):
        self.amplitude = amplitude
        self.summation_class = Superposition

    numeric = property(lambda self: self.derivation == Derivation.NUMERIC)
    analytic = property(lambda self: self.derivation == Derivation.ANALYTIC)
    variational = property(lambda self: self.derivation == Derivation.VARIATIONAL)
    bound = property(lambda self: self.binding == Binding.BOUND)
    free = property(lambda self: self.binding == Binding.FREE)

    @property
    def norm(self):
        return np.abs(self.amplitude) ** 2

    def normalized(self):
        return self / np.sqrt(self.norm)

    def __mul__(self, other):
        new = deepcopy(self)
        new.amplitude *= other
        return new

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self * (1 / other)

    @property
    def tuple(self):
        raise NotImplementedError

    def __hash__(self):
        return hash((self.__class__.__name__,) + self.tuple)

    def __eq__(self, other):
        return isinstance(other, self.__class__) and self.tuple == other.tuple

    def __lt__(self, other):
        return isinstance(other, self.__class__) and self.tuple < other.tuple

    def __gt__(self, other):
        return isinstance(other, self.__class__) and self.tuple > other.tuple

    def __le__(self, other):
        return isinstance(other, self.__class__) and self.tuple <= other.tuple

    def __ge__(self, other):
        return isinstance(other, self.__class__) and self.tuple >= other.tuple

    @property
    def ket(self):
        return f"|{self.__class__.__name__}>"

    def __str__(self):
        return self.ket

    __repr__ = __str__


class Superposition(Sum, QuantumState):
    """states/state.py:216-276"""

    def __init__(self, *states):
        amplitudes = collections.defaultdict(float)
        for s in states:
            amplitudes[s] += s.amplitude
        combined = []
        for s, amp in amplitudes.items():
            c = deepcopy(s)
            c.amplitude = amp
            combined.append(c)
        Sum.__init__(self, *combined)
        QuantumState.__init__(self, amplitude=np.sqrt(sum(s.norm for s in combined)))
        self.states = tuple(combined)

    @property
    def tuple(self):
        return sum((s.tuple for s in self), tuple())

    @property
    def ket(self):
        return " + ".join(s.ket for s in self)

    def normalized(self):
        return Superposition(*tuple(s / np.sqrt(self.norm) for s in self))

    @property
    def bound(self):
        return all(s.bound for s in self)

    @property
    def free(self):
        return all(s.free for s in self)


# ---------------------------------------------------------------------------------------------
# three-dimensional states
# ---------------------------------------------------------------------------------------------
class HydrogenBoundState(QuantumState):
    """states/three_d.py (constructor + radial_function :375-391)"""

    eigenvalues = Eigenvalues.DISCRETE
    binding = Binding.BOUND
    derivation = Derivation.ANALYTIC

    def __init__(self, n: int = 1, l: int = 0, m: int = 0, amplitude=1):
        if not (isinstance(n, (int, np.integer)) and n > 0):
            raise exceptions.IllegalQuantumState(f"n ({n}) must be an integer greater than zero")
        if not (0 <= l < n):
            raise exceptions.IllegalQuantumState(f"l ({l}) must be less than n ({n}) and greater than or equal to zero")
        if not (-l <= m <= l):
            raise exceptions.IllegalQuantumState(f"|m| (|{m}|) must be less than or equal to l ({l})")
        super().__init__(amplitude=amplitude)
        self.n, self.l, self.m = int(n), int(l), int(m)

    @property
    def energy(self):
        return -u.rydberg / (self.n ** 2)

    @property
    def spherical_harmonic(self):
        return SphericalHarmonic(l=self.l, m=self.m)

    @property
    def tuple(self):
        return self.n, self.l, self.m

    @property
    def ket(self):
        return f"|{self.n},{self.l},{self.m}>"

    def radial_function(self, r):
        n, l = self.n, self.l
        normalization = np.sqrt(((2 / (n * u.bohr_radius)) ** 3) * (special.factorial(n - l - 1) / (2 * n * special.factorial(n + l))))
        r_dep = np.exp(-r / (n * u.bohr_radius)) * ((2 * r / (n * u.bohr_radius)) ** l)
        lag_poly = special.eval_genlaguerre(n - l - 1, (2 * l) + 1, 2 * r / (n * u.bohr_radius))
        return self.amplitude * normalization * r_dep * lag_poly

    def __call__(self, r, theta, phi):
        return self.radial_function(r) * self.spherical_harmonic(theta, phi)


class NumericSphericalHarmonicState(QuantumState):
    """a radial eigenvector of the discretised H0 in channel l (states/three_d.py NumericSphericalHarmonicState)"""

    eigenvalues = Eigenvalues.DISCRETE
    derivation = Derivation.NUMERIC

    def __init__(self, *, g, l: int, m: int, energy: float, corresponding_analytic_state, binding, amplitude=1):
        super().__init__(amplitude=amplitude)
        self.g = g
        self.l, self.m = int(l), int(m)
        self.energy = energy
        self.analytic_state = corresponding_analytic_state
        self.binding = binding

    @property
    def n(self):
        return getattr(self.analytic_state, "n", None)

    @property
    def spherical_harmonic(self):
        return SphericalHarmonic(l=self.l, m=self.m)

    @property
    def tuple(self):
        return self.analytic_state.tuple

    @property
    def ket(self):
        return self.analytic_state.ket + "_n"

    def radial_function(self, r):
        return self.g


class HydrogenCoulombState(QuantumState):
    """label for a numeric free state (states/three_d.py HydrogenCoulombState); only identity/ordering is used here"""

    eigenvalues = Eigenvalues.CONTINUOUS
    binding = Binding.FREE
    derivation = Derivation.ANALYTIC

    def __init__(self, energy: float = 1 * u.eV, l: int = 0, amplitude=1):
        super().__init__(amplitude=amplitude)
        self.energy = energy
        self.l = int(l)
        self.m = 0

    @property
    def tuple(self):
        return self.energy, self.l, 0

    @property
    def ket(self):
        return f"|{self.energy / u.eV:.3f} eV,{self.l}>"


# ---------------------------------------------------------------------------------------------
# one-dimensional states
# ---------------------------------------------------------------------------------------------
class QHOState(QuantumState):
    """states/one_d.py:130-296"""

    smallest_n = 0
    eigenvalues = Eigenvalues.DISCRETE
    binding = Binding.BOUND
    derivation = Derivation.ANALYTIC

    def __init__(self, spring_constant: float, mass: float = u.electron_mass, n: int = 0, amplitude=1):
        self.n = n
        self.spring_constant = spring_constant
        self.mass = mass
        super().__init__(amplitude=amplitude)

    @classmethod
    def from_omega_and_mass(cls, omega, mass=u.electron_mass, n=0, amplitude=1):
        return cls(spring_constant=mass * (omega ** 2), mass=mass, n=n, amplitude=amplitude)

    @classmethod
    def from_potential(cls, potential, mass, n=0, amplitude=1):
        return cls(spring_constant=potential.spring_constant, mass=mass, n=n, amplitude=amplitude)

    @property
    def omega(self):
        return np.sqrt(self.spring_constant / self.mass)

    @property
    def energy(self):
        return u.hbar * self.omega * (self.n + 0.5)

    @property
    def period(self):
        return u.twopi / self.omega

    @property
    def tuple(self):
        return self.n, self.mass, self.omega

    @property
    def ket(self):
        return f"|{self.n}>"

    def __call__(self, x):
        norm = ((self.mass * self.omega / (u.pi * u.hbar)) ** (1 / 4)) / (np.float64(2 ** (self.n / 2)) * np.sqrt(np.float64(special.factorial(self.n))))
        exp = np.exp(-self.mass * self.omega * (x ** 2) / (2 * u.hbar))
        herm = special.hermite(self.n)(np.sqrt(self.mass * self.omega / u.hbar) * x)
        return self.amplitude * (norm * exp * herm).astype(np.complex128)


class GaussianWellState(QuantumState):
    """variational ground state of a Gaussian well (states/one_d.py:572-697)"""

    smallest_n = 0
    eigenvalues = Eigenvalues.DISCRETE
    binding = Binding.BOUND
    derivation = Derivation.VARIATIONAL

    def __init__(self, well_depth: float, well_width: float, mass: float, n: int = 0, well_center: float = 0, amplitude=1):
        self.well_depth = np.abs(well_depth)
        self.well_width = well_width
        self.well_center = well_center
        self.mass = mass
        self.n = n
        max_n = np.ceil(2 * np.sqrt(2 * mass * self.well_depth / (u.pi * (u.hbar ** 2))) * well_width) + 0.5
        if n > max_n:
            raise exceptions.IllegalQuantumState("Bound state energy must be less than zero")
        self.width = optimize.newton(
            lambda w: ((w ** 4) / (((well_width ** 2) + (w ** 2)) ** 1.5)) - ((u.hbar ** 2) / (4 * mass * well_width * self.well_depth)),
            well_width,
        )
        self.energy = -(well_width * self.well_depth / np.sqrt(((well_width ** 2) + (self.width ** 2))) + ((u.hbar ** 2) / (8 * mass * (self.width ** 2))))
        super().__init__(amplitude=amplitude)

    @classmethod
    def from_potential(cls, potential, mass, n=0, amplitude=1):
        return cls(potential.potential_extrema, potential.width, mass, n=n, well_center=potential.center, amplitude=amplitude)

    @property
    def tuple(self):
        return self.well_depth, self.well_width, self.mass, self.n

    @property
    def ket(self):
        return f"|{self.n}>"

    def __call__(self, x):
        return np.exp(-0.25 * (x / self.width) ** 2) / (np.sqrt(np.sqrt(u.twopi) * self.width))


class OneDPlaneWave(QuantumState):
    """label for free numeric 1-D states (states/one_d.py:20-128)"""

    eigenvalues = Eigenvalues.CONTINUOUS
    binding = Binding.FREE
    derivation = Derivation.ANALYTIC

    def __init__(self, wavenumber: float = u.twopi / u.nm, mass: float = u.electron_mass, amplitude=1):
        self.wavenumber = wavenumber
        self.mass = mass
        super().__init__(amplitude=amplitude)

    @classmethod
    def from_energy(cls, energy, k_sign=1, mass=u.electron_mass, amplitude=1):
        return cls(k_sign * np.sqrt(2 * mass * energy) / u.hbar, mass, amplitude=amplitude)

    @property
    def energy(self):
        return ((u.hbar * self.wavenumber) ** 2) / (2 * self.mass)

    @property
    def tuple(self):
        return self.wavenumber, self.mass

    def __call__(self, x):
        return np.exp(1j * self.wavenumber * x) / np.sqrt(u.twopi)


class NumericOneDState(QuantumState):
    """states/one_d.py:788-870"""

    eigenvalues = Eigenvalues.DISCRETE
    derivation = Derivation.NUMERIC

    def __init__(self, *, g, energy: float, corresponding_analytic_state, binding, amplitude=1):
        super().__init__(amplitude=amplitude)
        self.g = g
        self.energy = energy
        self.analytic_state = corresponding_analytic_state
        self.binding = binding

    @property
    def tuple(self):
        return self.analytic_state.tuple

    @property
    def ket(self):
        return self.analytic_state.ket + "_n"

    def __call__(self, z):
        return self.g
