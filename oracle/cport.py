"""ORACLE TEST INFRASTRUCTURE -- ctypes access to the C restatement (oracle/c/restate.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  ``build()`` compiles it with gcc via oracle/Makefile.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_c128p = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_long = ctypes.c_long


def build(with_ref=True):
    targets = ["all"] + (["ref"] if with_ref else [])
    subprocess.run(["make", "-C", _HERE, *targets], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build(with_ref=False)
        L = ctypes.CDLL(_LIB_PATH)
        L.ora_tdma.argtypes = [_long, _c128p, _c128p, _c128p, _c128p, _c128p, _c128p]
        L.ora_sh_len_so_steps.argtypes = [_long, _long, _c128p, _c128p, _f64p, _f64p, _f64p, _f64p, _long, _f64p, _f64p]
        L.ora_sh_vel_so_steps.argtypes = [_long, _long, _c128p, _c128p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _long, _f64p, _f64p]
        for fn in (L.ora_line_cn_len_steps, L.ora_line_so_len_steps):
            fn.argtypes = [_long, _long, _c128p, _c128p, _f64p, _f64p, _f64p, _long, _f64p, _f64p]
        L.ora_line_so_vel_steps.argtypes = [_long, _long, _c128p, _c128p, _f64p, ctypes.c_double, _f64p, _long, _f64p, _f64p]
        L.ora_norm.argtypes = [_long, _c128p, ctypes.c_double]
        L.ora_norm.restype = ctypes.c_double
        L.ora_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return int(lib().ora_num_threads())


def tdma(sub, diag, sup, d):
    n = len(d)
    x = np.empty(n, dtype=np.complex128)
    work = np.empty(2 * n, dtype=np.complex128)
    lib().ora_tdma(n, _c(sub), _c(diag), _c(sup), _c(d), x, work)
    return x


def sh_steps(problem, g=None, nsteps=None, start=0):
    """Advance ``g`` (default problem['g0']) by ``nsteps`` steps starting at step ``start``; returns new g."""
    kind = str(problem["kind"])
    g = _c(problem["g0"] if g is None else g).copy()
    L, R = g.shape
    taus = _f(problem["taus"])[start:]
    fields = _f(problem["fields"])[start:]
    n = len(taus) if nsteps is None else nsteps
    taus, fields = _f(taus[:n]), _f(fields[:n])
    if kind == "sh_len_so":
        lib().ora_sh_len_so_steps(L, R, g, _c(problem["h_diag"]), _f(problem["h_off"]), _f(problem["c_l"]), _f(problem["x_j"]), _f(problem["mask"]), n, taus, fields)
    elif kind == "sh_vel_so":
        lib().ora_sh_vel_so_steps(
            L, R, g, _c(problem["h_diag"]), _f(problem["h_off"]), _f(problem["c_l"]), _f(problem["f1_l"]), _f(problem["y_j"]), _f(problem["z_j"]),
            _f(problem["mask"]), n, taus, fields,
        )
    else:
        raise ValueError(kind)
    return g


def line_steps(problem, g=None, fields=None, nsteps=None):
    """``g``: (batch, Z) or (Z,); ``fields``: (nsteps, batch) or (nsteps,)."""
    kind = str(problem["kind"])
    g0 = _c(problem["g0"] if g is None else g)
    g = np.atleast_2d(g0).copy()
    B, Z = g.shape
    fields = _f(problem["fields"] if fields is None else fields)
    if fields.ndim == 1:
        fields = np.repeat(fields[:, None], B, axis=1)
    taus = _f(problem["taus"])
    n = len(taus) if nsteps is None else nsteps
    fields = _f(fields[:n])
    args = (B, Z, g, _c(problem["h_diag"]), _f(problem["h_off"]))
    if kind == "line_len_cn":
        lib().ora_line_cn_len_steps(*args, _f(problem["w_z"]), _f(problem["mask"]), n, taus, fields)
    elif kind == "line_len_so":
        lib().ora_line_so_len_steps(*args, _f(problem["w_z"]), _f(problem["mask"]), n, taus, fields)
    elif kind == "line_vel_so":
        lib().ora_line_so_vel_steps(*args, float(problem["v_pref"]), _f(problem["mask"]), n, taus, fields)
    else:
        raise ValueError(kind)
    return g.reshape(g0.shape)
