"""Exception hierarchy; mirrors ionization/exceptions.py:1-40 of the reference for the names the mesh path raises."""


class IonizationException(Exception):
    """Base class for all exceptions of this package (ionization/exceptions.py:1-4)."""


class InvalidPotentialParameter(IonizationException):
    pass


class InvalidMaskParameter(IonizationException):
    pass


class InvalidWrappingDirection(IonizationException):
    pass


class InvalidChoice(IonizationException):
    pass


class IllegalQuantumState(IonizationException):
    pass


class UnknownData(IonizationException):
    pass


class MissingDatastore(IonizationException):
    pass


class DuplicateDatastores(IonizationException):
    pass


# --- engine-specific (no counterpart in the reference: it has no native boundary that can fail) ---
class EngineError(IonizationException):
    """A C-ABI call returned a non-zero status; the message is ion_last_error()."""


class NativeLibraryMissing(EngineError):
    """libionization_b200.so has not been built; there is no CPU fallback."""


class NoCudaDevice(EngineError):
    """No usable CUDA device; there is no CPU fallback."""


class UnsupportedConfiguration(EngineError):
    """The requested (operators, evolution method, mesh) combination is not on the CUDA hot path."""
