"""ctypes binding of the C-ABI in ``include/ionization_b200.h``.

The shared library is built in-tree by ``ionization_b200.build`` (nvcc, sm_100a) into
``ionization_b200/_lib/libionization_b200.so``.  There is NO CPU fallback: if the library is
missing, or there is no CUDA device, every compute entry point raises.
"""
import ctypes
import os

import numpy as np

from . import exceptions

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ION_LIB") or os.path.join(_HERE, "_lib", "libionization_b200.so")  # ION_LIB: A/B experiments

# mirrors of the header's constants
ION_SH_LEN_SO = 0
ION_SH_VEL_SO = 1
ION_LINE_LEN_CN = 2
ION_LINE_LEN_SO = 3
ION_LINE_VEL_SO = 4
ION_SH_LEN_ADI = 5

OBS_NORM = 1
OBS_INNER_PRODUCTS = 2
OBS_NORM_BY_L = 4
OBS_R = 8
OBS_Z = 16
OBS_H0 = 32
OBS_NORM_WITHIN = 64

ION_ENODEVICE = -2
ABI_VERSION = 2  # include/ionization_b200.h: ION_ABI_VERSION
PEER_BLOB_BYTES = 96  # include/ionization_b200.h: ION_PEER_BLOB_BYTES

_lib = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int
_u32 = ctypes.c_uint32
_f64 = ctypes.c_double
_f64p = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); every symbol the header declares
SIGNATURES = {
    "ion_abi_version": (_i32, []),
    "ion_last_error": (ctypes.c_char_p, []),
    "ion_device_count": (_i32, []),
    "ion_tdma_c128": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32]),
    "ion_sim_create": (_i32, [_i32, _i64, _i64, _i64, _i32, ctypes.POINTER(_vp)]),
    "ion_sim_create_sharded": (_i32, [_i32, _i64, _i64, _i64, _i64, _i64, _i32, ctypes.POINTER(_vp)]),
    "ion_sim_destroy": (_i32, [_vp]),
    "ion_sim_set_stream": (_i32, [_vp, _vp]),
    "ion_sim_set_hamiltonian": (_i32, [_vp, _vp, _vp]),
    "ion_sim_set_len_coupling": (_i32, [_vp, _vp, _vp]),
    "ion_sim_set_vel_coupling": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "ion_sim_set_line_coupling": (_i32, [_vp, _vp, _f64]),
    "ion_sim_set_mask": (_i32, [_vp, _vp]),
    "ion_sim_set_observables": (_i32, [_vp, _f64, _vp, _i64, _vp, _vp, _i64, _vp]),
    "ion_sim_write_g": (_i32, [_vp, _vp]),
    "ion_sim_read_g": (_i32, [_vp, _vp]),
    "ion_sim_write_g_broadcast": (_i32, [_vp, _vp]),
    "ion_sim_step": (_i32, [_vp, _i64, _vp, _vp]),
    "ion_sim_observation_size": (_i64, [_vp, _u32]),
    "ion_sim_observe": (_i32, [_vp, _u32, _vp]),
    "ion_sim_run": (_i32, [_vp, _i64, _vp, _vp, _vp, _u32, _vp]),
    "ion_sim_synchronize": (_i32, [_vp]),
    "ion_sim_halo_buffer": (_i32, [_vp, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_i64)]),
    "ion_sim_export_peer": (_i32, [_vp, _vp, _i64]),
    "ion_sim_attach_peer": (_i32, [_vp, _i32, _vp, _i64, _i32]),
    "ion_sim_exchange_halos": (_i32, [_vp]),
    "ion_sim_prepare": (_i32, [_vp, _f64]),
    "ion_sim_reserve": (_i32, [_vp, _i64, _i64, _u32]),
    "ion_sim_halo_status": (_i32, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(_i32)]),
    "ion_sim_num_phases": (_i32, [_vp]),
    "ion_sim_phase_needs_halo": (_i32, [_vp, _i32]),
    "ion_sim_step_phase": (_i32, [_vp, _i32, _f64, _vp]),
    "ion_sim_device_psi": (_i32, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_i64)]),
    "ion_sim_launch_count": (_i64, [_vp]),
    "ion_num_kernel_kinds": (_i32, []),
    "ion_kernel_name": (ctypes.c_char_p, [_i32]),
    "ion_sim_profile": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "ion_fp64_peak": (_i32, [_i32, _f64p]),
    "ion_sinc_pulse_fields": (_i32, [_i32, _i32, _i64, _vp, _f64, _i64, _vp, _vp]),
}


def library_path() -> str:
    return LIB_PATH


def load():
    """Load the shared library (no device needed).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise exceptions.NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m ionization_b200.build` "
            "(nvcc, sm_100a).  ionization_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if os.environ.get("ION_LIB"):  # an older build loaded for an A/B experiment
                continue
            raise
        fn.restype = restype
        fn.argtypes = argtypes
    if not os.environ.get("ION_LIB") and int(lib.ion_abi_version()) != ABI_VERSION:
        raise exceptions.NativeLibraryMissing(
            f"{LIB_PATH} implements C-ABI version {int(lib.ion_abi_version())}, this package binds version {ABI_VERSION}: "
            "rebuild it with `python -m ionization_b200.build --force`"
        )
    _lib = lib
    return lib


def last_error() -> str:
    return load().ion_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (code {rc})" if what else f"{last_error()} (code {rc})"
    if rc == ION_ENODEVICE:
        raise exceptions.NoCudaDevice(msg)
    raise exceptions.EngineError(msg)


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def as_c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
