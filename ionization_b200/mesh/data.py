"""Time-indexed data of a MeshSimulation (ionization/mesh/data.py).

Same datastore classes, ``sim.data.<name>`` attributes, NaN-initialised arrays and exceptions as the reference.
The values come from the fused device reductions (kernel 4): each datastore declares which observables it needs
(``observables``) and receives them as a record dictionary -- either one record per data time while stepping
(``store``) or all of them at once after a device-resident run (``store_many``).
"""
import collections
import sys

import numpy as np

from .. import exceptions
from .. import _native as nat


class Data:
    """data.py:14-99"""

    def __init__(self, sim):
        self.sim = sim
        self.times = sim.data_times

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        datastore = DATA_NAME_TO_DATASTORE_TYPE.get(item)
        if datastore is None:
            raise exceptions.UnknownData(f"Couldn't find any data named '{item}' on {self.sim}. Ensure that the corresponding datastore is correctly implemented.")
        raise exceptions.MissingDatastore(f"Couldn't get data '{item}' for {self.sim} because it does not include a {datastore.__name__} datastore.")


DATA_NAME_TO_DATASTORE_TYPE = {}


class Datastore:
    observables = 0  # bit mask of engine observables this datastore consumes
    needs_mesh = False  # True: store() reads sim.mesh (host-side analysis); the run loop then stops at every data time

    def init(self, sim):
        self.sim = sim
        self.spec = sim.spec
        self.attach()

    def store(self, record, idx):
        """record: dict of observables at the current data time; idx: data time index"""
        raise NotImplementedError

    def attach(self):
        raise NotImplementedError

    def __repr__(self):
        return self.__class__.__name__


def _link(datastore_type, name):
    def getter(data):
        try:
            return getattr(data.sim.datastores_by_type[datastore_type], name)()
        except KeyError:
            raise exceptions.MissingDatastore(f"Couldn't get data {name} for {data.sim} because it does not include a {datastore_type.__name__} datastore.")

    setattr(Data, name, property(getter))


class Fields(Datastore):
    """data.py:152-186 -- host-side: E(t) and A(t) at the data times"""

    def init(self, sim):
        self.electric_field_amplitude = sim.get_blank_data()
        self.vector_potential_amplitude = sim.get_blank_data()
        super().init(sim)

    def store(self, record, idx):
        self.electric_field_amplitude[idx] = record["electric_field_amplitude"]
        self.vector_potential_amplitude[idx] = record["vector_potential_amplitude"]

    def attach(self):
        self.sim.data.electric_field_amplitude = self.electric_field_amplitude
        self.sim.data.vector_potential_amplitude = self.vector_potential_amplitude

    def __sizeof__(self):
        return self.electric_field_amplitude.nbytes + self.vector_potential_amplitude.nbytes + super().__sizeof__()


DATA_NAME_TO_DATASTORE_TYPE.update({"electric_field_amplitude": Fields, "vector_potential_amplitude": Fields})


class Norm(Datastore):
    """data.py:189-206"""

    observables = nat.OBS_NORM

    def init(self, sim):
        self.norm = sim.get_blank_data()
        super().init(sim)

    def store(self, record, idx):
        self.norm[idx] = record["norm"]

    def attach(self):
        self.sim.data.norm = self.norm

    def __sizeof__(self):
        return self.norm.nbytes + super().__sizeof__()


DATA_NAME_TO_DATASTORE_TYPE.update({"norm": Norm})


class InnerProducts(Datastore):
    """data.py:210-287"""

    observables = nat.OBS_INNER_PRODUCTS

    def init(self, sim):
        self.inner_products = {state: sim.get_blank_data(dtype=np.complex128) for state in sim.spec.test_states}
        super().init(sim)

    def store(self, record, idx):
        ips = record["inner_products"]
        for k, state in enumerate(self.spec.test_states):
            self.inner_products[state][idx] = ips[k]

    def state_overlaps(self):
        return {state: np.abs(ip) ** 2 for state, ip in self.inner_products.items()}

    def initial_state_overlap(self):
        return np.abs(self.inner_products[self.spec.initial_state]) ** 2

    def bound_state_overlap(self):
        return sum(ov for state, ov in self.state_overlaps().items() if state.bound)

    def free_state_overlap(self):
        return sum(ov for state, ov in self.state_overlaps().items() if state.free)

    def total_state_overlap(self):
        return sum(self.state_overlaps().values())

    def attach(self):
        self.sim.data.inner_products = self.inner_products
        self.sim.data.initial_state_inner_product = self.inner_products[self.spec.initial_state]

    def __sizeof__(self):
        return sum(ip.nbytes for ip in self.inner_products.values()) + sys.getsizeof(self.inner_products) + super().__sizeof__()


for _n in ("state_overlaps", "initial_state_overlap", "bound_state_overlap", "free_state_overlap", "total_state_overlap"):
    _link(InnerProducts, _n)
DATA_NAME_TO_DATASTORE_TYPE.update(
    {n: InnerProducts for n in ("inner_products", "initial_state_inner_product", "state_overlaps", "initial_state_overlap", "bound_state_overlap",
                                "free_state_overlap", "total_state_overlap")}
)


class InternalEnergyExpectationValue(Datastore):
    """data.py:290-310"""

    observables = nat.OBS_H0

    def init(self, sim):
        self.internal_energy_expectation_value = sim.get_blank_data()
        super().init(sim)

    def store(self, record, idx):
        self.internal_energy_expectation_value[idx] = record["internal_energy"]

    def attach(self):
        self.sim.data.internal_energy_expectation_value = self.internal_energy_expectation_value


DATA_NAME_TO_DATASTORE_TYPE.update({"internal_energy_expectation_value": InternalEnergyExpectationValue})


class TotalEnergyExpectationValue(Datastore):
    """data.py:317-339: <H0> + <H_int(t)>; for the length gauge H_int = E * (-q) z (mesh_operators.py:1008-1035, :320-327)"""

    observables = nat.OBS_H0 | nat.OBS_Z

    def init(self, sim):
        self.total_energy_expectation_value = sim.get_blank_data()
        super().init(sim)

    def store(self, record, idx):
        self.total_energy_expectation_value[idx] = record["total_energy"]

    def attach(self):
        self.sim.data.total_energy_expectation_value = self.total_energy_expectation_value


DATA_NAME_TO_DATASTORE_TYPE.update({"total_energy_expectation_value": TotalEnergyExpectationValue})


class ZExpectationValue(Datastore):
    """data.py:346-377"""

    observables = nat.OBS_Z

    def init(self, sim):
        self.z_expectation_value = sim.get_blank_data()
        super().init(sim)

    def store(self, record, idx):
        self.z_expectation_value[idx] = record["z"]

    def attach(self):
        self.sim.data.z_expectation_value = self.z_expectation_value

    def z_dipole_moment_expectation_value(self):
        return self.spec.test_charge * self.z_expectation_value


_link(ZExpectationValue, "z_dipole_moment_expectation_value")
DATA_NAME_TO_DATASTORE_TYPE.update({"z_expectation_value": ZExpectationValue, "z_dipole_moment_expectation_value": ZExpectationValue})


class RExpectationValue(Datastore):
    """data.py:380-399"""

    observables = nat.OBS_R

    def init(self, sim):
        self.r_expectation_value = sim.get_blank_data()
        super().init(sim)

    def store(self, record, idx):
        self.r_expectation_value[idx] = record["r"]

    def attach(self):
        self.sim.data.r_expectation_value = self.r_expectation_value


DATA_NAME_TO_DATASTORE_TYPE.update({"r_expectation_value": RExpectationValue})


class NormWithinRadius(Datastore):
    """data.py:402-435"""

    observables = nat.OBS_NORM_WITHIN

    def __init__(self, radii=()):
        self.radii = tuple(sorted(radii))

    def init(self, sim):
        self.norm_within_radius = {r: sim.get_blank_data() for r in self.radii}
        super().init(sim)

    def store(self, record, idx):
        for k, r in enumerate(self.radii):
            self.norm_within_radius[r][idx] = record["norm_within_radius"][k]

    def attach(self):
        self.sim.data.norm_within_radius = self.norm_within_radius


DATA_NAME_TO_DATASTORE_TYPE.update({"norm_within_radius": NormWithinRadius})


class NormBySphericalHarmonic(Datastore):
    """data.py:438-464.  (In the reference ``init`` reads ``self.spec`` before it is set, so the datastore cannot be
    attached there; the intended behaviour -- ``sim.data.norm_by_sph_harm[SphericalHarmonic(l, 0)]`` -- is provided.)"""

    observables = nat.OBS_NORM_BY_L

    def init(self, sim):
        self.norm_by_l = {sph_harm: sim.get_blank_data() for sph_harm in sim.spec.spherical_harmonics}
        super().init(sim)

    def store(self, record, idx):
        for sph_harm, l_norm in zip(self.spec.spherical_harmonics, record["norm_by_l"]):
            self.norm_by_l[sph_harm][idx] = l_norm

    def attach(self):
        self.sim.data.norm_by_sph_harm = self.norm_by_l
        self.sim.data.norm_by_l = self.norm_by_l


DATA_NAME_TO_DATASTORE_TYPE.update({"norm_by_l": NormBySphericalHarmonic, "norm_by_sph_harm": NormBySphericalHarmonic})

class DirectionalRadialProbabilityCurrent(Datastore):
    """data.py:467-531: the radial probability current as a function of radius, integrated over the upper (z > 0) and the lower
    hemisphere.  Host-side analysis on the (r, theta) reconstruction of the synchronised wavefunction at the data times."""

    needs_mesh = True

    def init(self, sim):
        self.radial_probability_current__pos_z = np.zeros((sim.data_time_steps, sim.spec.r_points), dtype=np.float64) * np.nan
        self.radial_probability_current__neg_z = np.zeros((sim.data_time_steps, sim.spec.r_points), dtype=np.float64) * np.nan
        theta = sim.mesh.theta_calc
        self.d_theta = np.abs(theta[1] - theta[0])
        self.sin_theta = np.sin(theta)
        self.mask = theta <= np.pi / 2
        super().init(sim)

    def store(self, record, idx):
        mesh = self.sim.mesh
        radial_current_density = mesh.get_radial_probability_current_density_mesh__spatial()
        integrand = radial_current_density * self.sin_theta * self.d_theta * (2 * np.pi)  # sin(theta) d_theta, two pi from phi
        self.radial_probability_current__pos_z[idx] = np.sum(integrand[:, self.mask], axis=1) * (mesh.r ** 2)
        self.radial_probability_current__neg_z[idx] = np.sum(integrand[:, ~self.mask], axis=1) * (mesh.r ** 2)

    def attach(self):
        self.sim.data.radial_probability_current__pos_z = self.radial_probability_current__pos_z
        self.sim.data.radial_probability_current__neg_z = self.radial_probability_current__neg_z

    def radial_probability_current__total(self):
        return self.radial_probability_current__pos_z + self.radial_probability_current__neg_z

    def __sizeof__(self):
        return self.radial_probability_current__pos_z.nbytes + self.radial_probability_current__neg_z.nbytes + super().__sizeof__()


_link(DirectionalRadialProbabilityCurrent, "radial_probability_current__total")
DATA_NAME_TO_DATASTORE_TYPE.update({
    "radial_probability_current__pos_z": DirectionalRadialProbabilityCurrent,
    "radial_probability_current__neg_z": DirectionalRadialProbabilityCurrent,
    "radial_probability_current__total": DirectionalRadialProbabilityCurrent,
})

DATASTORE_TYPE_TO_DATA_NAMES = collections.defaultdict(set)
for _data_name, _datastore_type in DATA_NAME_TO_DATASTORE_TYPE.items():
    DATASTORE_TYPE_TO_DATA_NAMES[_datastore_type].add(_data_name)

DEFAULT_DATASTORE_TYPES = (Fields, Norm, InnerProducts)
