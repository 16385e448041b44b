"""ionization_b200 -- B200-native implementation of the mesh time-evolution hot path of JoshKarpel/ionization.

Public surface (mirrors ``ionization`` for this path; see DESIGN.md / INTEGRATION.md):

    ionization_b200.mesh        specifications, simulations, operators, evolution methods, datastores
    ionization_b200.potentials  pulses, windows, static potentials, masks (host-side inputs)
    ionization_b200.states      hydrogen / 1-D states (host-side inputs)
    ionization_b200.engine      the CUDA engine behind the C-ABI (include/ionization_b200.h)
    ionization_b200.parallel    one-process-per-GPU ensembles and l-block sharding (torch.distributed)
    ionization_b200.scan        ParameterScan container (ionization/analysis.py) and the scan runner sharded over GPUs

The compute path is hand-written sm_100a CUDA behind a C-ABI shared library; there is no CPU fallback.
"""
from .version import __version__  # noqa: F401
from . import exceptions, units  # noqa: F401
from . import potentials, states  # noqa: F401
from . import mesh  # noqa: F401
from .core import Gauge  # noqa: F401
