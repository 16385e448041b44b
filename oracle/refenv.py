"""ORACLE TEST INFRASTRUCTURE -- not product code; container-only.

Imports the UNMODIFIED reference package from /root/reference under the
``simulacra`` shim (oracle/shim).  The reference compiles ``cy.pyx`` at import
through pyximport into ``ionization/.pyxbld`` (ionization/__init__.py:16-22),
which must be writable, so the package is mirrored into a scratch directory
OUTSIDE the repo (never into the repo: reference sources are not copied here).

``/root/reference`` does not exist on the GPU box, so nothing that runs there
(``-m gpu`` tests, smoke(), bench.py) may import this module.  It is used only
by ``oracle/make_golden.py`` and by the container-only pinning tests.
"""
import os
import shutil
import sys
import tempfile
import warnings

REFERENCE_ROOT = os.environ.get("IONIZATION_REFERENCE_ROOT", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ionization"))


def import_reference():
    """Return the imported reference ``ionization`` module (and install the shim)."""
    if "ionization" in sys.modules and getattr(sys.modules["ionization"], "_b200_oracle_ref", False):
        return sys.modules["ionization"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    scratch = os.path.join(tempfile.gettempdir(), "ionization_b200_oracle_ref")
    pkg = os.path.join(scratch, "ionization")
    if not os.path.isdir(pkg):
        os.makedirs(scratch, exist_ok=True)
        shutil.copytree(os.path.join(REFERENCE_ROOT, "ionization"), pkg)

    if _SHIM_DIR not in sys.path:
        sys.path.insert(0, _SHIM_DIR)
    if scratch not in sys.path:
        sys.path.insert(0, scratch)

    warnings.filterwarnings("ignore")
    import simulacra  # noqa: F401  (installs compat patches + stubs)
    import ionization

    ionization._b200_oracle_ref = True
    return ionization
