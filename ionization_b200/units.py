"""SI unit constants (CODATA-2014), the values the reference takes from ``simulacra.units`` (an un-vendored
dependency: requirements.txt:9 of the reference).  They reproduce the reference's six known answers
(dev/meshes/mesh_refactoring_helper.py:204-251) to the 12 printed digits; see tests/test_host_inputs.py.
"""
import numpy as np

pi = np.pi
twopi = 2 * np.pi
e = np.e
alpha = 7.2973525664e-3

m = 1.0
cm = 1e-2
mm = 1e-3
um = 1e-6
nm = 1e-9
pm = 1e-12
angstrom = 1e-10
per_nm = 1 / nm

s = 1.0
msec = 1e-3
usec = 1e-6
nsec = 1e-9
psec = 1e-12
fsec = 1e-15
asec = 1e-18

Hz = 1.0
kHz = 1e3
MHz = 1e6
GHz = 1e9
THz = 1e12

kg = 1.0
J = 1.0
N = 1.0
W = 1.0
TW = 1e12
C = 1.0
V = 1.0

c = 299792458.0
mu_0 = pi * 4e-7
epsilon_0 = 1 / (mu_0 * c ** 2)
coulomb_constant = 1 / (4 * pi * epsilon_0)
h = 6.626070040e-34
hbar = h / twopi
proton_charge = 1.6021766208e-19
electron_charge = -proton_charge
eV = proton_charge
keV = 1e3 * eV
MeV = 1e6 * eV
proton_mass = 1.672621898e-27
electron_mass = 9.10938356e-31
electron_mass_reduced = electron_mass * proton_mass / (electron_mass + proton_mass)
bohr_radius = 5.2917721067e-11
per_bohr_radius = 1 / bohr_radius
rydberg = 13.605693009 * eV
hartree = 2 * rydberg
atomic_time = hbar / hartree
atomic_electric_field = coulomb_constant * proton_charge / bohr_radius ** 2
atomic_electric_potential = coulomb_constant * proton_charge / bohr_radius
atomic_momentum = hbar / bohr_radius
atomic_force = hartree / bohr_radius
atomic_velocity = alpha * c
atomic_angular_frequency = 1 / atomic_time
atomic_electric_dipole_moment = proton_charge * bohr_radius
atomic_intensity = 0.5 * epsilon_0 * c * atomic_electric_field ** 2
Jcm2 = J / cm ** 2
Wcm2 = W / cm ** 2
TWcm2 = TW / cm ** 2
V_per_m = 1.0
N_per_m = 1.0
deg = pi / 180
rad = 1.0


class Unit:
    pass


def get_unit_value_and_latex(x):
    if isinstance(x, str):
        return globals().get(x, 1), x
    return (1 if x is None else x), str(x)
