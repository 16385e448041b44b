"""Python face of the CUDA engine: a thin, typed wrapper over the C-ABI handle (include/ionization_b200.h).

Nothing here computes: every method forwards host buffers to the shared library, which does the
host<->device copies and launches the sm_100a kernels.  The reference-facing API (specifications,
simulations, datastores) lives in ``ionization_b200.mesh`` and drives this class.
"""
import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _native as nat
from . import exceptions

PROGRAMS = {
    "sh_len_so": nat.ION_SH_LEN_SO,
    "sh_vel_so": nat.ION_SH_VEL_SO,
    "line_len_cn": nat.ION_LINE_LEN_CN,
    "line_len_so": nat.ION_LINE_LEN_SO,
    "line_vel_so": nat.ION_LINE_VEL_SO,
    "sh_len_adi": nat.ION_SH_LEN_ADI,
}


def device_count() -> int:
    return int(nat.load().ion_device_count())


def fp64_peak(device: int = 0) -> float:
    """measured FP64 pipe peak of the device in thread-level FMAs per second (x2 = FLOP/s)"""
    v = ctypes.c_double(0.0)
    nat.check(nat.load().ion_fp64_peak(int(device), ctypes.byref(v)), "ion_fp64_peak")
    return float(v.value)


def tdma(matrix, d, device: int = 0):
    """Drop-in for ``ionization.cy.tdma(matrix, d)`` (cy.pyx:9-50): x = matrix^-1 d, no pivoting.

    ``matrix`` is a scipy ``dia_matrix`` with offsets (-1, 0, 1) exactly as the reference passes it
    (cy.pyx:20-22), or a tuple ``(sub, diag, sup)``.  ``d`` may be ``(n,)`` or ``(batch, n)``.
    Inputs are not modified; a new array is returned.
    """
    if isinstance(matrix, (tuple, list)):
        sub, diag, sup = matrix
    else:
        offsets = list(matrix.offsets)
        data = matrix.data
        sub = data[offsets.index(-1)][:-1]  # cy.pyx:20: subdiagonal[1:] = matrix.data[0]
        diag = data[offsets.index(0)]  # cy.pyx:21
        sup = data[offsets.index(1)][1:]  # cy.pyx:22
    d = np.asarray(d)
    single = d.ndim == 1
    d2 = nat.as_c128(np.atleast_2d(d))
    batch, n = d2.shape
    sub = nat.as_c128(np.broadcast_to(np.atleast_2d(sub), (batch, max(n - 1, 0))))
    sup = nat.as_c128(np.broadcast_to(np.atleast_2d(sup), (batch, max(n - 1, 0))))
    diag = nat.as_c128(np.broadcast_to(np.atleast_2d(diag), (batch, n)))
    x = np.empty_like(d2)
    lib = nat.load()
    nat.check(lib.ion_tdma_c128(nat.ptr(sub), nat.ptr(diag), nat.ptr(sup), nat.ptr(d2), nat.ptr(x), n, batch, device), "ion_tdma_c128")
    return x[0] if single else x


class DeviceSimulation:
    """``batch`` simulations on one mesh, wavefunction resident on one GPU."""

    def __init__(self, program, L: int, R: int, batch: int = 1, device: int = 0, L_total: Optional[int] = None, l_begin: int = 0):
        self._lib = nat.load()
        self._h = ctypes.c_void_p()
        self.program = PROGRAMS[program] if isinstance(program, str) else int(program)
        self.L, self.R, self.batch, self.device = int(L), int(R), int(batch), int(device)
        self.L_total = int(L if L_total is None else L_total)
        self.l_begin = int(l_begin)
        self.n_states = 0
        self.n_radii = 0
        sharded = self.L != self.L_total
        self.g_lo = 1 if (sharded and self.l_begin > 0) else 0
        self.g_hi = 1 if (sharded and self.l_begin + self.L < self.L_total) else 0
        nat.check(
            self._lib.ion_sim_create_sharded(self.program, self.L_total, self.l_begin, self.L, self.R, self.batch, self.device, ctypes.byref(self._h)),
            "ion_sim_create",
        )

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ion_sim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- set-up -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        nat.check(self._lib.ion_sim_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "ion_sim_set_stream")

    def set_hamiltonian(self, h_diag, h_off):
        """h_diag: [L (+ ghost channels of an l-block shard), R]; h_off: [R-1]"""
        h_diag = nat.as_c128(np.asarray(h_diag).reshape(self.L + self.g_lo + self.g_hi, self.R))
        h_off = nat.as_f64(h_off)
        if h_off.shape != (self.R - 1,):
            raise exceptions.EngineError(f"h_off must have shape ({self.R - 1},), got {h_off.shape}")
        nat.check(self._lib.ion_sim_set_hamiltonian(self._h, nat.ptr(h_diag), nat.ptr(h_off)), "ion_sim_set_hamiltonian")

    def set_len_coupling(self, c_l, x_j):
        c_l, x_j = nat.as_f64(c_l), nat.as_f64(x_j)
        self._shape(c_l, (self.L_total - 1,), "c_l")
        self._shape(x_j, (self.R,), "x_j")
        nat.check(self._lib.ion_sim_set_len_coupling(self._h, nat.ptr(c_l), nat.ptr(x_j)), "ion_sim_set_len_coupling")

    def set_vel_coupling(self, c_l, f1_l, y_j, z_j):
        c_l, f1_l, y_j, z_j = map(nat.as_f64, (c_l, f1_l, y_j, z_j))
        self._shape(c_l, (self.L_total - 1,), "c_l")
        self._shape(f1_l, (self.L_total - 1,), "f1_l")
        self._shape(y_j, (self.R,), "y_j")
        self._shape(z_j, (self.R - 1,), "z_j")
        nat.check(self._lib.ion_sim_set_vel_coupling(self._h, nat.ptr(c_l), nat.ptr(f1_l), nat.ptr(y_j), nat.ptr(z_j)), "ion_sim_set_vel_coupling")

    def set_line_coupling(self, w_z=None, v_pref: float = 0.0):
        if w_z is not None:
            w_z = nat.as_f64(w_z)
            self._shape(w_z, (self.R,), "w_z")
        nat.check(self._lib.ion_sim_set_line_coupling(self._h, nat.ptr(w_z), float(v_pref)), "ion_sim_set_line_coupling")

    def set_mask(self, mask):
        if mask is not None:
            mask = nat.as_f64(np.broadcast_to(np.asarray(mask, dtype=np.float64), (self.R,)))
        nat.check(self._lib.ion_sim_set_mask(self._h, nat.ptr(mask)), "ion_sim_set_mask")

    def set_observables(self, inner_product_multiplier: float, r=None, state_l: Sequence[int] = (), state_rows=None, radii: Sequence[float] = ()):
        r = None if r is None else nat.as_f64(r)
        state_l = np.ascontiguousarray(state_l, dtype=np.int64)
        n_states = len(state_l)
        rows = None
        if n_states:
            rows = nat.as_c128(state_rows)
            self._shape(rows, (n_states, self.R), "state_rows")
        radii = nat.as_f64(radii)
        nat.check(
            self._lib.ion_sim_set_observables(
                self._h, float(inner_product_multiplier), nat.ptr(r), n_states, nat.ptr(state_l) if n_states else None, nat.ptr(rows), len(radii),
                nat.ptr(radii) if len(radii) else None,
            ),
            "ion_sim_set_observables",
        )
        self.n_states, self.n_radii = n_states, len(radii)

    @staticmethod
    def _shape(a, shape, name):
        if a.shape != tuple(shape):
            raise exceptions.EngineError(f"{name} must have shape {tuple(shape)}, got {a.shape}")

    # -- wavefunction -----------------------------------------------------------------------
    @property
    def g_shape(self):
        return (self.batch, self.L, self.R)

    def write_g(self, g):
        g = nat.as_c128(g)
        if g.size != self.batch * self.L * self.R:
            raise exceptions.EngineError(f"g must have {self.batch}x{self.L}x{self.R} elements, got shape {g.shape}")
        nat.check(self._lib.ion_sim_write_g(self._h, nat.ptr(g)), "ion_sim_write_g")
        self.synchronize()  # g may be a temporary

    def write_g_broadcast(self, g):
        """the same wavefunction ``g`` [L, R] for every member of the ensemble (replicated on the device)"""
        g = nat.as_c128(g)
        if g.size != self.L * self.R:
            raise exceptions.EngineError(f"g must have {self.L}x{self.R} elements, got shape {g.shape}")
        nat.check(self._lib.ion_sim_write_g_broadcast(self._h, nat.ptr(g)), "ion_sim_write_g_broadcast")
        self.synchronize()

    def read_g(self, out=None):
        if out is None:
            out = np.empty(self.g_shape, dtype=np.complex128)
        nat.check(self._lib.ion_sim_read_g(self._h, nat.ptr(out)), "ion_sim_read_g")
        return out

    # -- evolution --------------------------------------------------------------------------
    def _scalars(self, taus, fields):
        taus = nat.as_f64(np.atleast_1d(taus))
        n = len(taus)
        fields = nat.as_f64(fields)
        if fields.size == n and self.batch != 1:
            fields = nat.as_f64(np.repeat(fields.reshape(n, 1), self.batch, axis=1))
        if fields.size != n * self.batch:
            raise exceptions.EngineError(f"fields must have shape ({n}, {self.batch}), got {fields.shape}")
        return n, taus, fields

    def step(self, taus, fields):
        """advance len(taus) steps; asynchronous"""
        n, taus, fields = self._scalars(taus, fields)
        nat.check(self._lib.ion_sim_step(self._h, n, nat.ptr(taus), nat.ptr(fields)), "ion_sim_step")

    def observation_size(self, what: int) -> int:
        return int(self._lib.ion_sim_observation_size(self._h, what))

    def observe(self, what: int):
        out = np.empty((self.batch, self.observation_size(what)), dtype=np.float64)
        nat.check(self._lib.ion_sim_observe(self._h, what, nat.ptr(out)), "ion_sim_observe")
        return out

    def run(self, taus, fields, observe_mask=None, what: int = 0):
        """the device-resident loop: returns records [n_observed, batch, observation_size]"""
        n, taus, fields = self._scalars(taus, fields)
        if observe_mask is None:
            observe_mask = np.zeros(n, dtype=np.uint8)
        observe_mask = np.ascontiguousarray(observe_mask, dtype=np.uint8)
        n_obs = int(np.count_nonzero(observe_mask))
        out = np.empty((n_obs, self.batch, self.observation_size(what)), dtype=np.float64)
        nat.check(
            self._lib.ion_sim_run(self._h, n, nat.ptr(taus), nat.ptr(fields), nat.ptr(observe_mask), what, nat.ptr(out) if out.size else None),
            "ion_sim_run",
        )
        return out

    def synchronize(self):
        nat.check(self._lib.ion_sim_synchronize(self._h), "ion_sim_synchronize")

    # -- l-block shards: one step = a few pair-local phases with halo exchanges in between -------
    @property
    def num_phases(self) -> int:
        return int(self._lib.ion_sim_num_phases(self._h))

    def phase_needs_halo(self, phase: int) -> bool:
        return bool(self._lib.ion_sim_phase_needs_halo(self._h, phase))

    def step_phase(self, phase: int, tau: float, field):
        field = nat.as_f64(np.broadcast_to(np.asarray(field, dtype=np.float64), (self.batch,)))
        nat.check(self._lib.ion_sim_step_phase(self._h, phase, float(tau), nat.ptr(field)), "ion_sim_step_phase")

    # -- peer-memory halo exchange inside the engine (csrc/halo.cuh) ------------------------------
    def export_peer(self) -> bytes:
        """the blob this shard's neighbours need (CUDA IPC handles + ghost offsets)"""
        buf = ctypes.create_string_buffer(nat.PEER_BLOB_BYTES)
        nat.check(self._lib.ion_sim_export_peer(self._h, buf, nat.PEER_BLOB_BYTES), "ion_sim_export_peer")
        return bytes(buf.raw)

    def attach_peer(self, side: int, blob: bytes, same_process: bool = False):
        """side 0: the lower neighbour's blob, side 1: the upper neighbour's"""
        buf = ctypes.create_string_buffer(blob, len(blob))
        nat.check(self._lib.ion_sim_attach_peer(self._h, int(side), buf, len(blob), 1 if same_process else 0), "ion_sim_attach_peer")

    def exchange_halos(self):
        nat.check(self._lib.ion_sim_exchange_halos(self._h), "ion_sim_exchange_halos")

    def prepare(self, tau: float):
        nat.check(self._lib.ion_sim_prepare(self._h, float(tau)), "ion_sim_prepare")

    def reserve(self, n_steps: int, n_records: int = 0, what: int = 0):
        """size the buffers of a later step / run call now (linked shards must not allocate between hand-shakes)"""
        nat.check(self._lib.ion_sim_reserve(self._h, int(n_steps), int(n_records), int(what)), "ion_sim_reserve")

    def halo_status(self):
        n, ab = ctypes.c_int64(0), ctypes.c_int32(0)
        nat.check(self._lib.ion_sim_halo_status(self._h, ctypes.byref(n), ctypes.byref(ab)), "ion_sim_halo_status")
        return int(n.value), bool(ab.value)

    def halo_buffer(self, which: int):
        """(device pointer, bytes) of a boundary-channel buffer: 0 send-to-lower, 1 send-to-upper, 2 recv-from-lower,
        3 recv-from-upper (None when that neighbour does not exist)"""
        p = ctypes.c_void_p()
        nbytes = ctypes.c_int64()
        nat.check(self._lib.ion_sim_halo_buffer(self._h, which, ctypes.byref(p), ctypes.byref(nbytes)), "ion_sim_halo_buffer")
        return p.value, nbytes.value

    # -- measurement ------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self._lib.ion_sim_launch_count(self._h))

    def profile(self, taus, fields):
        """per-kernel-kind (milliseconds, launches) over len(taus) steps, CUDA events around every launch"""
        n, taus, fields = self._scalars(taus, fields)
        k = int(self._lib.ion_num_kernel_kinds())
        ms = np.zeros(k, dtype=np.float64)
        cnt = np.zeros(k, dtype=np.int64)
        nat.check(self._lib.ion_sim_profile(self._h, n, nat.ptr(taus), nat.ptr(fields), nat.ptr(ms), nat.ptr(cnt)), "ion_sim_profile")
        names = [self._lib.ion_kernel_name(i).decode() for i in range(k)]
        return {nm: (float(m), int(c)) for nm, m, c in zip(names, ms, cnt) if c}

    def device_psi(self):
        p = ctypes.c_void_p()
        nbytes = ctypes.c_int64()
        nat.check(self._lib.ion_sim_device_psi(self._h, ctypes.byref(p), ctypes.byref(nbytes)), "ion_sim_device_psi")
        return p.value, nbytes.value

    # -- convenience: build from a dict of hot-path inputs (the layout of tests/golden/*.npz) -----
    @classmethod
    def from_problem(cls, problem, batch: int = 1, device: int = 0, with_states: bool = True, radii=()):
        kind = str(problem["kind"])
        if kind.startswith("sh"):
            L, R = int(problem["L"]), int(problem["R"])
            sim = cls(kind, L, R, batch=batch, device=device)
            sim.set_hamiltonian(problem["h_diag"], problem["h_off"])
            if kind == "sh_vel_so":
                sim.set_vel_coupling(problem["c_l"], problem["f1_l"], problem["y_j"], problem["z_j"])
            else:
                sim.set_len_coupling(problem["c_l"], problem["x_j"])
            sim.set_mask(problem["mask"])
            sl = problem["state_l"] if with_states and "state_l" in problem else ()
            sr = problem["state_rows"] if with_states and "state_rows" in problem else None
            sim.set_observables(float(problem["delta_r"]), problem["r"], sl, sr, radii)
            g0 = np.asarray(problem["g0"], dtype=np.complex128).reshape(L, R)
        else:
            R = int(problem["Z"])
            sim = cls(kind, 1, R, batch=batch, device=device)
            sim.set_hamiltonian(np.asarray(problem["h_diag"]).reshape(1, R), problem["h_off"])
            sim.set_line_coupling(problem["w_z"], float(problem["v_pref"]))
            sim.set_mask(problem["mask"])
            rows = problem["state_rows"] if with_states and "state_rows" in problem else None
            sl = np.zeros(len(rows), dtype=np.int64) if rows is not None else ()
            sim.set_observables(float(problem["delta_z"]), problem["z"], sl, rows, radii)
            g0 = np.asarray(problem["g0"], dtype=np.complex128).reshape(1, R)
        sim.write_g_broadcast(g0)
        return sim
