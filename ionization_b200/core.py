"""Small shared definitions (reference: ionization/core.py:127, mesh/meshes.py:44-53, mesh/mesh_operators.py:25-27)."""
import enum


class StrEnum(str, enum.Enum):
    def __str__(self):
        return self.value


class Gauge(StrEnum):
    LENGTH = "LEN"
    VELOCITY = "VEL"


class WrappingDirection(StrEnum):
    Z = "z"
    R = "r"
    L = "l"


class KineticEnergyDerivation(StrEnum):
    HAMILTONIAN = "hamiltonian"
    LAGRANGIAN = "lagrangian"


class Status(StrEnum):
    INITIALIZED = "initialized"
    RUNNING = "running"
    FINISHED = "finished"
    PAUSED = "paused"
    ERROR = "error"
