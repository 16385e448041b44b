"""Host-side physics inputs of the hot path: static potentials, electric fields (pulses + time windows) and masks.

These are evaluated on the host -- once at set-up (potentials, masks) or once per time step into a scalar
(fields) -- and handed to the CUDA engine as plain vectors/scalars, so they are not on the measured path
("INPUT to hot path").  The classes keep the reference's names, constructor arguments
and formulas (ionization/potentials/*.py) so existing scripts keep working; any object with the same methods
(e.g. the reference's own pulse objects) can be used instead -- the mesh layer only duck-types:

    potential(r=..., test_charge=...)            static potential energy on the mesh
    pulse.get_electric_field_amplitude(t)        E(t)
    pulse.get_vector_potential_amplitude_numeric(times)   A(times[-1]) = -integral E dt
    mask(r=...)                                  mask values in [0, 1]
"""
import functools

import numpy as np
import scipy.optimize as optim

from . import exceptions
from . import units as u


# ---------------------------------------------------------------------------------------------
# summation algebra (ionization/summables.py:6-77)
# ---------------------------------------------------------------------------------------------
class Summand:
    summation_class = None

    def __iter__(self):
        yield self

    def __add__(self, other):
        return (self.summation_class or Sum)(*self, *other)

    def __str__(self):
        return self.__class__.__name__

    __repr__ = __str__


class Sum(Summand):
    def __init__(self, *summands):
        self.summands = tuple(summands)

    def __iter__(self):
        yield from self.summands

    def __getitem__(self, item):
        return self.summands[item]

    def __add__(self, other):
        return self.__class__(*self, *other)

    def __call__(self, *args, **kwargs):
        return sum(x(*args, **kwargs) for x in self.summands)

    def __str__(self):
        return "(" + " + ".join(str(s) for s in self.summands) + ")"

    __repr__ = __str__


# ---------------------------------------------------------------------------------------------
# numerical integration rules the reference uses for A(t) and fluence
# ---------------------------------------------------------------------------------------------
def _basic_simpson(y, x, start, stop):
    """composite Simpson on samples start..stop (inclusive count odd), non-uniform spacing."""
    s0 = slice(start, stop, 2)
    s1 = slice(start + 1, stop + 1, 2)
    s2 = slice(start + 2, stop + 2, 2)
    h = np.diff(x)
    h0, h1 = h[s0], h[s1]
    hsum, hprod, h0divh1 = h0 + h1, h0 * h1, h0 / h1
    return np.sum(hsum / 6.0 * (y[s0] * (2 - 1.0 / h0divh1) + y[s1] * hsum * hsum / hprod + y[s2] * (2 - h0divh1)))


def simps(y, x):
    """``scipy.integrate.simps(y, x)`` as shipped up to scipy 1.10 (default ``even='avg'``), which is what the
    reference's A(t) is defined by (potentials/pulses.py:58-77).  For an even number of samples: the average of
    (Simpson on the first N-1 samples + trapezoid on the last interval) and (trapezoid on the first interval +
    Simpson on the last N-1 samples).  Newer scipy's ``simpson`` treats the even case differently."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    N = len(y)
    if N < 2:
        return 0.0
    if N == 2:
        return 0.5 * (x[1] - x[0]) * (y[0] + y[1])
    if N % 2 == 1:
        return float(_basic_simpson(y, x, 0, N - 2))
    first = 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2]) + _basic_simpson(y, x, 0, N - 3)
    last = 0.5 * (x[1] - x[0]) * (y[1] + y[0]) + _basic_simpson(y, x, 1, N - 2)
    return float(0.5 * (first + last))


def cumtrapz0(y, x):
    """``scipy.integrate.cumtrapz(y, x, initial=0)``"""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros(len(y))
    out[1:] = np.cumsum(0.5 * np.diff(x) * (y[1:] + y[:-1]))
    return out


def trapz(y, x):
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return float(np.sum(0.5 * np.diff(x) * (y[1:] + y[:-1])))


_RULES = {"simps": simps, "trapz": trapz}


# ---------------------------------------------------------------------------------------------
# potential energies (ionization/potentials/potential.py, static.py)
# ---------------------------------------------------------------------------------------------
class PotentialEnergy(Summand):
    def __init__(self, *args, **kwargs):
        self.summation_class = PotentialEnergySum


class PotentialEnergySum(Sum, PotentialEnergy):
    """potentials/potential.py:22-62"""

    def __init__(self, *potentials):
        Sum.__init__(self, *potentials)
        self.summation_class = PotentialEnergySum

    @property
    def potentials(self):
        return self.summands

    @property
    def window(self):
        return self.summands[0].window

    def get_electric_field_amplitude(self, t):
        return sum(x.get_electric_field_amplitude(t) for x in self.summands)

    def get_vector_potential_amplitude_numeric(self, times, rule="simps"):
        return sum(x.get_vector_potential_amplitude_numeric(times, rule=rule) for x in self.summands)

    def get_electric_field_integral_numeric_cumulative(self, times):
        return sum(x.get_electric_field_integral_numeric_cumulative(times) for x in self.summands)

    def get_vector_potential_amplitude_numeric_cumulative(self, times):
        return sum(x.get_vector_potential_amplitude_numeric_cumulative(times) for x in self.summands)

    def get_fluence_numeric(self, times, rule="simps"):
        return u.epsilon_0 * u.c * _RULES[rule](np.abs(self.get_electric_field_amplitude(times)) ** 2, times)


class NoPotentialEnergy(PotentialEnergy):
    def __call__(self, *, r, **kwargs):
        return np.zeros_like(r)


class CoulombPotential(PotentialEnergy):
    """potentials/static.py:12-38"""

    def __init__(self, charge: float = 1 * u.proton_charge):
        super().__init__()
        self.charge = charge

    def __call__(self, *, r, test_charge, **kwargs):
        return u.coulomb_constant * self.charge * test_charge / r

    def __repr__(self):
        return f"CoulombPotential(charge = {self.charge})"


class SoftCoulombPotential(PotentialEnergy):
    """potentials/static.py (soft-core Coulomb): k q Q / sqrt(r^2 + softening^2)"""

    def __init__(self, charge: float = 1 * u.proton_charge, softening_distance: float = 0.05 * u.bohr_radius):
        super().__init__()
        self.charge = charge
        self.softening_distance = softening_distance

    def __call__(self, *, r, test_charge, **kwargs):
        return u.coulomb_constant * self.charge * test_charge / np.sqrt(r ** 2 + self.softening_distance ** 2)


class HarmonicOscillator(PotentialEnergy):
    """potentials/static.py:113-217"""

    def __init__(self, spring_constant: float = 4.20521 * u.N / u.m, center: float = 0 * u.nm, cutoff_distance=None):
        super().__init__()
        self.spring_constant = spring_constant
        self.center = center
        self.cutoff_distance = cutoff_distance

    @classmethod
    def from_frequency_and_mass(cls, omega: float = 1.5192675e15 * u.Hz, mass: float = u.electron_mass, **kwargs):
        return cls(spring_constant=mass * (omega ** 2), **kwargs)

    @classmethod
    def from_ground_state_energy_and_mass(cls, ground_state_energy: float = 0.5 * u.eV, mass: float = u.electron_mass, **kwargs):
        return cls.from_frequency_and_mass(omega=2 * ground_state_energy / u.hbar, mass=mass, **kwargs)

    @classmethod
    def from_energy_spacing_and_mass(cls, energy_spacing: float = 1 * u.eV, mass: float = u.electron_mass, **kwargs):
        return cls.from_frequency_and_mass(omega=energy_spacing / u.hbar, mass=mass, **kwargs)

    def __call__(self, *, r, **kwargs):
        centered_r = r - self.center
        inside = 0.5 * self.spring_constant * (centered_r ** 2)
        if self.cutoff_distance is not None:
            outside = 0.5 * self.spring_constant * (self.cutoff_distance ** 2)
            return np.where(np.less_equal(np.abs(centered_r), self.cutoff_distance), inside, outside)
        return inside

    def omega(self, mass: float) -> float:
        return np.sqrt(self.spring_constant / mass)


class GaussianPotential(PotentialEnergy):
    """potentials/static.py:327-360"""

    def __init__(self, potential_extrema: float = -1 * u.eV, width: float = 1 * u.bohr_radius, center: float = 0):
        super().__init__()
        self.potential_extrema = potential_extrema
        self.width = width
        self.center = center

    def __call__(self, *, r, **kwargs):
        centered_r = r - self.center
        return self.potential_extrema * np.exp(-0.5 * ((centered_r / self.width) ** 2))


class ImaginaryGaussianRing(PotentialEnergy):
    """Complex absorbing ring (potentials/imaginary.py): -i * decay_energy * exp(-((r - center)/width)^2 / 2).
    It only makes the Crank-Nicolson diagonal complex, which the engine supports."""

    def __init__(self, center: float = 20 * u.bohr_radius, width: float = 2 * u.bohr_radius, decay_time: float = 100 * u.asec):
        super().__init__()
        self.center = center
        self.width = width
        self.decay_time = decay_time
        self.prefactor = -1j * u.hbar / self.decay_time

    def __call__(self, *, r, **kwargs):
        return self.prefactor * np.exp(-0.5 * (((r - self.center) / self.width) ** 2))


# ---------------------------------------------------------------------------------------------
# time windows (ionization/potentials/windows.py)
# ---------------------------------------------------------------------------------------------
class TimeWindow(Summand):
    def __init__(self):
        self.summation_class = TimeWindowSum


class TimeWindowSum(Sum, TimeWindow):
    def __call__(self, *args, **kwargs):
        return functools.reduce(lambda a, b: a * b, (x(*args, **kwargs) for x in self.summands))


class NoTimeWindow(TimeWindow):
    def __call__(self, t):
        return 1


class RectangularWindow(TimeWindow):
    """windows.py: 1 for start_time <= t <= end_time, else 0"""

    def __init__(self, start_time: float = 0 * u.asec, end_time: float = 50 * u.asec):
        super().__init__()
        self.start_time = start_time
        self.end_time = end_time

    def __call__(self, t):
        cond = np.greater_equal(t, self.start_time) * np.less_equal(t, self.end_time)
        return np.where(cond, 1.0, 0.0)


class LogisticWindow(TimeWindow):
    """windows.py:129-168"""

    def __init__(self, *, window_time: float, window_width: float, window_center: float = 0 * u.asec):
        super().__init__()
        self.window_time = window_time
        self.window_width = window_width
        self.window_center = window_center

    def __call__(self, t):
        tau = np.array(t) - self.window_center
        return (1 / (1 + np.exp(-(tau + self.window_time) / self.window_width))) - (
            1 / (1 + np.exp(-(tau - self.window_time) / self.window_width))
        )

    def __repr__(self):
        return f"LogisticWindow(window_time = {self.window_time}, window_width = {self.window_width}, window_center = {self.window_center})"


# ---------------------------------------------------------------------------------------------
# electric fields (ionization/potentials/pulses.py)
# ---------------------------------------------------------------------------------------------
class ElectricPotential(PotentialEnergy):
    pass


class UniformLinearlyPolarizedElectricPotential(ElectricPotential):
    """potentials/pulses.py:23-97"""

    def __init__(self, window: TimeWindow = None):
        super().__init__()
        self.window = window if window is not None else NoTimeWindow()

    def get_electric_field_amplitude(self, t):
        return self.window(t)

    def __call__(self, *, t, z, test_charge, **kwargs):
        return -z * test_charge * self.get_electric_field_amplitude(t)

    def get_electric_field_integral_numeric(self, times, rule: str = "simps"):
        return _RULES[rule](self.get_electric_field_amplitude(times), times)

    def get_vector_potential_amplitude_numeric(self, times, rule: str = "simps"):
        """A(times[-1]) = -integral of E over ``times`` (pulses.py:75-77)."""
        return -self.get_electric_field_integral_numeric(times, rule=rule)

    def get_electric_field_integral_numeric_cumulative(self, times):
        return cumtrapz0(self.get_electric_field_amplitude(times), times)

    def get_vector_potential_amplitude_numeric_cumulative(self, times):
        return -self.get_electric_field_integral_numeric_cumulative(times)

    def get_fluence_numeric(self, times, rule: str = "simps"):
        return u.epsilon_0 * u.c * _RULES[rule](np.abs(self.get_electric_field_amplitude(times)) ** 2, times)


class NoElectricPotential(UniformLinearlyPolarizedElectricPotential):
    def get_electric_field_amplitude(self, t):
        return np.zeros(np.shape(t)) * super().get_electric_field_amplitude(t)


class Rectangle(UniformLinearlyPolarizedElectricPotential):
    """pulses.py:111-155"""

    def __init__(self, start_time: float = 0 * u.asec, end_time: float = 50 * u.asec, amplitude: float = 1 * u.atomic_electric_field, **kwargs):
        if start_time >= end_time:
            raise exceptions.InvalidPotentialParameter("end_time must be later than start_time")
        super().__init__(**kwargs)
        self.start_time = start_time
        self.end_time = end_time
        self.amplitude = amplitude

    def get_electric_field_amplitude(self, t):
        cond = np.greater_equal(t, self.start_time) * np.less_equal(t, self.end_time)
        return np.where(cond, np.ones(np.shape(t)), np.zeros(np.shape(t))) * self.amplitude * super().get_electric_field_amplitude(t)


class SineWave(UniformLinearlyPolarizedElectricPotential):
    """pulses.py:215-426"""

    def __init__(self, omega: float, amplitude: float = 1 * u.atomic_electric_field, phase: float = 0, **kwargs):
        if omega <= 0:
            raise exceptions.InvalidPotentialParameter("omega must be positive")
        super().__init__(**kwargs)
        self.omega = omega
        self.phase = phase % u.twopi
        self.amplitude = amplitude

    @classmethod
    def from_frequency(cls, frequency, amplitude=1 * u.atomic_electric_field, phase=0, **kwargs):
        return cls(frequency * u.twopi, amplitude=amplitude, phase=phase, **kwargs)

    @classmethod
    def from_period(cls, period, amplitude=1 * u.atomic_electric_field, phase=0, **kwargs):
        return cls.from_frequency(1 / period, amplitude=amplitude, phase=phase, **kwargs)

    @classmethod
    def from_photon_energy(cls, photon_energy, amplitude=1 * u.atomic_electric_field, phase=0, **kwargs):
        return cls(photon_energy / u.hbar, amplitude=amplitude, phase=phase, **kwargs)

    @property
    def frequency(self):
        return self.omega / u.twopi

    @property
    def period(self):
        return 1 / self.frequency

    @property
    def photon_energy(self):
        return u.hbar * self.omega

    def get_electric_field_amplitude(self, t):
        return np.sin((self.omega * t) + self.phase) * self.amplitude * super().get_electric_field_amplitude(t)


DEFAULT_PULSE_WIDTH = 200 * u.asec
DEFAULT_FLUENCE = 1 * u.Jcm2
DEFAULT_PHASE = 0
DEFAULT_OMEGA_MIN = u.twopi * 30 * u.THz
DEFAULT_OMEGA_CARRIER = u.twopi * 2530 * u.THz
DEFAULT_PULSE_CENTER = 0 * u.asec


def sinc(x):
    """sin(x)/x (pulses.py:594-596)"""
    return np.sinc(x / u.pi)


class SincPulse(UniformLinearlyPolarizedElectricPotential):
    """pulses.py:599-1013 (constructor :646-695, field :929-940)"""

    def __init__(self, pulse_width=DEFAULT_PULSE_WIDTH, fluence=DEFAULT_FLUENCE, phase=DEFAULT_PHASE, pulse_center=DEFAULT_PULSE_CENTER,
                 omega_min=DEFAULT_OMEGA_MIN, **kwargs):
        if pulse_width <= 0:
            raise exceptions.InvalidPotentialParameter("pulse width must be positive")
        if fluence < 0:
            raise exceptions.InvalidPotentialParameter("fluence must be non-negative")
        if omega_min <= 0:
            raise exceptions.InvalidPotentialParameter("omega_min must be positive")
        super().__init__(**kwargs)
        self.omega_min = omega_min
        self.pulse_width = pulse_width
        self.phase = phase % u.twopi
        self.fluence = fluence
        self.pulse_center = pulse_center
        self.delta_omega = u.twopi / self.pulse_width
        self.omega_max = self.omega_min + self.delta_omega
        self.omega_carrier = (self.omega_min + self.omega_max) / 2
        self.amplitude_omega = np.sqrt(self.fluence / (2 * u.epsilon_0 * u.c * self.delta_omega))
        self.amplitude = np.sqrt(self.fluence * self.delta_omega / (u.pi * u.epsilon_0 * u.c))

    @classmethod
    def from_omega_min(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    @classmethod
    def from_omega_carrier(cls, pulse_width=DEFAULT_PULSE_WIDTH, fluence=DEFAULT_FLUENCE, phase=DEFAULT_PHASE, pulse_center=DEFAULT_PULSE_CENTER,
                           omega_carrier=DEFAULT_OMEGA_CARRIER, **kwargs):
        delta_omega = u.twopi / pulse_width
        return cls(pulse_width=pulse_width, fluence=fluence, phase=phase, pulse_center=pulse_center, omega_min=omega_carrier - delta_omega / 2, **kwargs)

    @property
    def photon_energy_carrier(self):
        return u.hbar * self.omega_carrier

    @property
    def number_of_cycles(self):
        return self.omega_carrier / self.delta_omega

    def get_electric_field_envelope(self, t):
        tau = np.array(t) - self.pulse_center
        return sinc(self.delta_omega * tau / 2)

    def get_electric_field_amplitude(self, t):
        tau = np.array(t) - self.pulse_center
        amp = self.get_electric_field_envelope(t) * np.cos((self.omega_carrier * tau) + self.phase)
        return amp * self.amplitude * super().get_electric_field_amplitude(t)

    def __repr__(self):
        return f"SincPulse(pulse_width = {self.pulse_width}, pulse_center = {self.pulse_center}, fluence = {self.fluence}, phase = {self.phase}, window = {self.window})"


class GaussianPulse(UniformLinearlyPolarizedElectricPotential):
    """pulses.py:1016-1372"""

    def __init__(self, pulse_width=DEFAULT_PULSE_WIDTH, omega_carrier=DEFAULT_OMEGA_CARRIER, fluence=DEFAULT_FLUENCE, phase=DEFAULT_PHASE,
                 pulse_center=DEFAULT_PULSE_CENTER, **kwargs):
        if pulse_width <= 0:
            raise exceptions.InvalidPotentialParameter("pulse width must be positive")
        if fluence < 0:
            raise exceptions.InvalidPotentialParameter("fluence must be non-negative")
        if omega_carrier < 0:
            raise exceptions.InvalidPotentialParameter("omega_carrier must be non-negative")
        super().__init__(**kwargs)
        self.omega_carrier = omega_carrier
        self.pulse_width = pulse_width
        self.phase = phase % u.twopi
        self.fluence = fluence
        self.pulse_center = pulse_center
        self.delta_omega = 1 / pulse_width
        self.amplitude = np.sqrt(2 * self.fluence / (np.sqrt(u.pi) * u.epsilon_0 * u.c * self.pulse_width))
        self.amplitude_omega = self.amplitude * self.pulse_width / 2

    @classmethod
    def from_omega_carrier(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    @classmethod
    def from_number_of_cycles(cls, pulse_width=DEFAULT_PULSE_WIDTH, number_of_cycles=3, number_of_pulse_widths=3, fluence=DEFAULT_FLUENCE,
                              phase=DEFAULT_PHASE, pulse_center=DEFAULT_PULSE_CENTER, **kwargs):
        omega_carrier = u.pi * number_of_cycles / (number_of_pulse_widths * pulse_width)
        pulse = cls(pulse_width=pulse_width, omega_carrier=omega_carrier, fluence=fluence, phase=phase, pulse_center=pulse_center, **kwargs)
        pulse.number_of_cycles = number_of_cycles
        pulse.number_of_pulse_widths = number_of_pulse_widths
        return pulse

    def get_electric_field_envelope(self, t):
        tau = np.array(t) - self.pulse_center
        return np.exp(-0.5 * ((tau / self.pulse_width) ** 2))

    def get_electric_field_amplitude(self, t):
        tau = t - self.pulse_center
        amp = self.get_electric_field_envelope(t) * np.cos((self.omega_carrier * tau) + self.phase)
        return amp * self.amplitude * super().get_electric_field_amplitude(t)


def DC_correct_electric_potential(electric_potential, times):
    """pulses.py:1961-2000: add a constant (windowed) field so the net field integral over ``times`` vanishes."""

    def func_to_minimize(amp, original_pulse):
        test = original_pulse + Rectangle(start_time=times[0], end_time=times[-1], amplitude=amp, window=electric_potential.window)
        return np.abs(test.get_electric_field_integral_numeric_cumulative(times)[-1])

    correction_amp = optim.minimize_scalar(func_to_minimize, args=(electric_potential,)).x
    correction = Rectangle(start_time=times[0], end_time=times[-1], amplitude=correction_amp, window=electric_potential.window)
    return electric_potential + correction


class FluenceCorrector(UniformLinearlyPolarizedElectricPotential):
    """pulses.py:2003-2030"""

    def __init__(self, electric_potential, times, target_fluence):
        self.electric_potential = electric_potential
        self.target_fluence = target_fluence
        fluence = electric_potential.get_fluence_numeric(times)
        self.amplitude_correction_ratio = np.sqrt(target_fluence / fluence)
        super().__init__()

    def get_electric_field_amplitude(self, t):
        return self.electric_potential.get_electric_field_amplitude(t) * self.amplitude_correction_ratio


# ---------------------------------------------------------------------------------------------
# masks (ionization/potentials/masks.py)
# ---------------------------------------------------------------------------------------------
class Mask(Summand):
    def __init__(self):
        self.summation_class = MaskSum


class MaskSum(Sum, Mask):
    """masks multiply (masks.py:24-28)"""

    def __init__(self, *masks):
        Sum.__init__(self, *masks)
        self.summation_class = MaskSum

    def __call__(self, *args, **kwargs):
        return functools.reduce(lambda a, b: a * b, (x(*args, **kwargs) for x in self.summands))


class NoMask(Mask):
    def __call__(self, *args, **kwargs):
        return 1


class RadialCosineMask(Mask):
    """masks.py:37-89: 1 inside inner_radius, |cos(pi/2 (r - ri)/(ro - ri))|^(1/smoothness) on the ramp, 0 outside."""

    def __init__(self, inner_radius: float = 50 * u.bohr_radius, outer_radius: float = 100 * u.bohr_radius, smoothness: float = 8):
        if inner_radius < 0 or outer_radius < 0:
            raise exceptions.InvalidMaskParameter("inner and outer radius must be non-negative")
        if inner_radius >= outer_radius:
            raise exceptions.InvalidMaskParameter("outer radius must be larger than inner radius")
        if smoothness < 1:
            raise exceptions.InvalidMaskParameter("smoothness must be greater than 1")
        super().__init__()
        self.inner_radius = inner_radius
        self.outer_radius = outer_radius
        self.smoothness = smoothness

    def __call__(self, *, r, **kwargs):
        r = np.asarray(r)
        return np.where(
            np.greater_equal(r, self.inner_radius) * np.less(r, self.outer_radius),
            np.abs(np.cos(0.5 * u.pi * (r - self.inner_radius) / np.abs(self.outer_radius - self.inner_radius))) ** (1 / self.smoothness),
            np.where(np.greater_equal(r, self.outer_radius), 0, 1),
        )

    def __repr__(self):
        return f"RadialCosineMask(inner_radius = {self.inner_radius}, outer_radius = {self.outer_radius}, smoothness = {self.smoothness})"
