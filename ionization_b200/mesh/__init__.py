"""``ionization_b200.mesh`` -- drop-in for the mesh time-evolution path of ``ionization.mesh``
(specifications, simulations, operators, evolution methods, datastores), backed by the CUDA engine."""
from ..core import Gauge, KineticEnergyDerivation, WrappingDirection  # noqa: F401
from .data import (  # noqa: F401
    Data,
    Datastore,
    Fields,
    InnerProducts,
    InternalEnergyExpectationValue,
    Norm,
    NormBySphericalHarmonic,
    NormWithinRadius,
    RExpectationValue,
    TotalEnergyExpectationValue,
    ZExpectationValue,
    DirectionalRadialProbabilityCurrent,
    DEFAULT_DATASTORE_TYPES,
    DATA_NAME_TO_DATASTORE_TYPE,
    DATASTORE_TYPE_TO_DATA_NAMES,
)
from .evolution_methods import AlternatingDirectionImplicit, EvolutionMethod, SplitInteractionOperator  # noqa: F401
from .meshes import LineMesh, QuantumMesh, SphericalHarmonicMesh  # noqa: F401
from .operators import (  # noqa: F401
    LineLengthGaugeOperators,
    LineVelocityGaugeOperators,
    MeshOperators,
    SphericalHarmonicLengthGaugeOperators,
    SphericalHarmonicVelocityGaugeOperators,
)
from .sims import (  # noqa: F401
    LineSpecification,
    MeshSimulation,
    MeshSpecification,
    SphericalHarmonicSimulation,
    SphericalHarmonicSpecification,
)
from .snapshots import Snapshot, SphericalHarmonicSnapshot  # noqa: F401
from .ensemble import MeshEnsemble, run_ensemble  # noqa: F401
