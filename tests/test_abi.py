"""The C-ABI shared library loads without a GPU and exports every symbol include/ionization_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from ionization_b200 import _native


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ionization_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ion_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_native.LIB_PATH):
        from ionization_b200 import build

        build.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    # the python binding covers exactly the header
    assert sorted(_native.SIGNATURES) == names


def test_library_reports_abi_version_and_fails_cleanly_without_device():
    lib = _native.load()
    assert lib.ion_abi_version() == _native.ABI_VERSION
    if lib.ion_device_count() == 0:
        h = ctypes.c_void_p()
        rc = lib.ion_sim_create(0, 4, 16, 1, 0, ctypes.byref(h))
        assert rc == _native.ION_ENODEVICE
        assert b"CPU" in lib.ion_last_error() or b"device" in lib.ion_last_error()
