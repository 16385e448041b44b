"""ORACLE TEST INFRASTRUCTURE -- not product code.

Minimal stand-in for the reference's un-vendored ``simulacra`` dependency
(requirements.txt:9) so that ``/root/reference/ionization`` can be imported in
this container to pin the oracle and to generate golden fixtures.  It provides
only what ``ionization.mesh`` touches on the hot path: Specification/Simulation
bases, Status, Info, a few ``utils`` helpers, ``math.SphericalHarmonic`` and
stubs for plotting modules that are absent here (matplotlib, cycler).

It also restores APIs that newer numpy/scipy removed and the 2019-era reference
still calls: np.NaN/np.Inf, scipy.integrate.simps (the OLD ``even='avg'``
algorithm of scipy<=1.10 -- reference call site potentials/pulses.py:58-77),
cumtrapz, trapz.
"""
import collections
import datetime
import enum
import functools
import os
import pickle
import sys
import time
import types
import uuid

import numpy as np

from . import units  # noqa: F401


# --------------------------------------------------------------------------
# stubs for absent plotting packages
# --------------------------------------------------------------------------
class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any


class _Any(metaclass=_AnyMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Any()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any


def _stub(name):
    mod = _StubModule(name)
    mod.__path__ = []
    sys.modules[name] = mod
    return mod


for _n in (
    "matplotlib",
    "matplotlib.pyplot",
    "matplotlib.colors",
    "matplotlib.animation",
    "mpl_toolkits",
    "mpl_toolkits.axes_grid1",
    "cycler",
):
    if _n not in sys.modules:
        try:
            __import__(_n)
        except Exception:
            _stub(_n)
vis = _stub("simulacra.vis")

# --------------------------------------------------------------------------
# numpy / scipy compatibility for the 2019-era reference
# --------------------------------------------------------------------------
if not hasattr(np, "NaN"):
    np.NaN = np.nan
if not hasattr(np, "Inf"):
    np.Inf = np.inf

import scipy  # noqa: E402
import scipy.integrate as _integ  # noqa: E402
import scipy.interpolate as _interp  # noqa: E402


def _tupleset(t, i, value):
    lst = list(t)
    lst[i] = value
    return tuple(lst)


def _basic_simps(y, start, stop, x, dx, axis):
    nd = len(y.shape)
    if start is None:
        start = 0
    step = 2
    slice_all = (slice(None),) * nd
    s0 = _tupleset(slice_all, axis, slice(start, stop, step))
    s1 = _tupleset(slice_all, axis, slice(start + 1, stop + 1, step))
    s2 = _tupleset(slice_all, axis, slice(start + 2, stop + 2, step))
    if x is None:
        return np.sum(dx / 3.0 * (y[s0] + 4 * y[s1] + y[s2]), axis=axis)
    h = np.diff(x, axis=axis)
    h0 = h[s0]
    h1 = h[s1]
    hsum = h0 + h1
    hprod = h0 * h1
    h0divh1 = h0 / h1
    return np.sum(
        hsum / 6.0 * (y[s0] * (2 - 1.0 / h0divh1) + y[s1] * hsum * hsum / hprod + y[s2] * (2 - h0divh1)),
        axis=axis,
    )


def simps(y, x=None, dx=1, axis=-1, even="avg"):
    """Composite Simpson with the pre-1.11 scipy treatment of even sample counts."""
    y = np.asarray(y)
    nd = len(y.shape)
    N = y.shape[axis]
    last_dx = dx
    first_dx = dx
    if x is not None:
        x = np.asarray(x)
    if N % 2 == 0:
        val = 0.0
        result = 0.0
        slice1 = (slice(None),) * nd
        slice2 = (slice(None),) * nd
        if even in ("avg", "first"):
            slice1 = _tupleset(slice1, axis, -1)
            slice2 = _tupleset(slice2, axis, -2)
            if x is not None:
                last_dx = x[slice1] - x[slice2]
            val += 0.5 * last_dx * (y[slice1] + y[slice2])
            result = _basic_simps(y, 0, N - 3, x, dx, axis)
        if even in ("avg", "last"):
            slice1 = _tupleset(slice1, axis, 0)
            slice2 = _tupleset(slice2, axis, 1)
            if x is not None:
                first_dx = x[slice2] - x[slice1]
            val += 0.5 * first_dx * (y[slice2] + y[slice1])
            result += _basic_simps(y, 1, N - 2, x, dx, axis)
        if even == "avg":
            val /= 2.0
            result /= 2.0
        return result + val
    return _basic_simps(y, 0, N - 2, x, dx, axis)


if not hasattr(_integ, "simps"):
    _integ.simps = simps
if not hasattr(_integ, "cumtrapz"):
    _integ.cumtrapz = _integ.cumulative_trapezoid
if not hasattr(_integ, "trapz"):
    _integ.trapz = _integ.trapezoid
import scipy.special as _special  # noqa: E402

if not hasattr(_special, "sph_harm"):  # removed upstream; old signature sph_harm(m, n, azimuth, polar) (reference: meshes.py:1173, :1483)
    _special.sph_harm = lambda m, n, theta, phi: _special.sph_harm_y(n, m, phi, theta)
if not hasattr(scipy, "interp"):
    scipy.interp = _interp


# --------------------------------------------------------------------------
# Specification / Simulation bases
# --------------------------------------------------------------------------
class Status(enum.Enum):
    INITIALIZED = "initialized"
    RUNNING = "running"
    FINISHED = "finished"
    PAUSED = "paused"
    ERROR = "error"


class Info:
    def __init__(self, *, header):
        self.header = header
        self.children = collections.OrderedDict()

    def add_field(self, name, value):
        self.children[name] = value

    def add_fields(self, name_value_pairs):
        for n, v in name_value_pairs:
            self.add_field(n, v)

    def add_info(self, info):
        self.children[id(info)] = info

    def add_infos(self, *infos):
        for i in infos:
            self.add_info(i)

    def __str__(self):
        return self.header

    def log(self, *a, **k):
        pass


class Beet:
    def __init__(self, name, file_name=None):
        self.name = str(name)
        self.file_name = file_name or self.name
        self.uuid = uuid.uuid4()

    def __eq__(self, other):
        return isinstance(other, self.__class__) and self.uuid == other.uuid

    def __hash__(self):
        return hash(self.uuid)

    def __str__(self):
        return f"{self.__class__.__name__}({self.name})"

    __repr__ = __str__

    def save(self, target_dir=None, file_extension="beet", **kwargs):
        path = os.path.join(str(target_dir or os.getcwd()), f"{self.file_name}.{file_extension}")
        with open(path, "wb") as f:
            pickle.dump(self, f)
        return path

    @classmethod
    def load(cls, path):
        with open(str(path), "rb") as f:
            return pickle.load(f)

    def info(self):
        return Info(header=str(self))


class Specification(Beet):
    simulation_type = None

    def __init__(self, name, file_name=None, **kwargs):
        super().__init__(name, file_name=file_name)
        self._extra_attr_keys = list(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to_sim(self):
        return self.simulation_type(self)

    def save(self, target_dir=None, file_extension="spec", **kwargs):
        return super().save(target_dir, file_extension)


class Simulation(Beet):
    def __init__(self, spec):
        super().__init__(spec.name, file_name=spec.file_name)
        self.spec = spec
        self.status = Status.INITIALIZED

    def save(self, target_dir=None, file_extension="sim", **kwargs):
        return super().save(target_dir, file_extension)

    def run(self):
        raise NotImplementedError


# --------------------------------------------------------------------------
# simulacra.utils
# --------------------------------------------------------------------------
utils = types.ModuleType("simulacra.utils")
sys.modules["simulacra.utils"] = utils


class StrEnum(str, enum.Enum):
    def __str__(self):
        return self.value


def memoize(func):
    memo = {}

    @functools.wraps(func)
    def memoizer(*args, **kwargs):
        key = (args, tuple(sorted(kwargs.items())))
        try:
            return memo[key]
        except KeyError:
            memo[key] = func(*args, **kwargs)
            return memo[key]

    return memoizer


def watched_memoize(watcher):
    class Watcher:
        def __init__(self, func):
            self.func = func
            self.cached = None
            self.watched = None

        def __call__(self, *args):
            w = watcher(args[0])
            if self.watched != w or self.cached is None:
                self.cached = self.func(*args)
                self.watched = w
            return self.cached

        def __get__(self, instance, owner):
            return functools.partial(self.__call__, instance)

    return Watcher


NearestEntry = collections.namedtuple("NearestEntry", ["index", "value", "target"])


def find_nearest_entry(array, target):
    array = np.asarray(array)
    i = int(np.argmin(np.abs(array - target)))
    return NearestEntry(i, array[i], target)


def bytes_to_str(n):
    return f"{n} B"


class BlockTimer:
    def __enter__(self):
        self._w0 = time.perf_counter()
        self._p0 = time.process_time()
        return self

    def __exit__(self, *exc):
        self.wall_time_elapsed = datetime.timedelta(seconds=time.perf_counter() - self._w0)
        self.proc_time_elapsed = datetime.timedelta(seconds=time.process_time() - self._p0)


def dict_to_arrays(d):
    return np.array(list(d.keys())), np.array(list(d.values()))


for _k, _v in dict(
    StrEnum=StrEnum,
    memoize=memoize,
    watched_memoize=watched_memoize,
    cached_property=functools.cached_property,
    find_nearest_entry=find_nearest_entry,
    bytes_to_str=bytes_to_str,
    BlockTimer=BlockTimer,
    dict_to_arrays=dict_to_arrays,
).items():
    setattr(utils, _k, _v)

# --------------------------------------------------------------------------
# simulacra.math
# --------------------------------------------------------------------------
math = types.ModuleType("simulacra.math")
sys.modules["simulacra.math"] = math


class SphericalHarmonic:
    def __init__(self, l=0, m=0):
        self.l = l
        self.m = m

    def __hash__(self):
        return hash((self.l, self.m))

    def __eq__(self, other):
        return isinstance(other, SphericalHarmonic) and (self.l, self.m) == (other.l, other.m)

    def __lt__(self, other):
        return (self.l, self.m) < (other.l, other.m)

    def __repr__(self):
        return f"SphericalHarmonic(l={self.l}, m={self.m})"

    def __call__(self, theta, phi=0):
        import scipy.special as sp

        return sp.sph_harm_y(self.l, self.m, theta, phi)


def complex_nquad(*a, **k):
    raise NotImplementedError


math.SphericalHarmonic = SphericalHarmonic
math.complex_nquad = complex_nquad
