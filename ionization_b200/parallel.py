"""One process per GPU: scan-ensemble sharding and l-block sharding of one large simulation.

* Scan ensembles shard as INDEPENDENT simulations: ``shard_range`` splits the members over the ranks, every rank
  runs its own batched ``DeviceSimulation`` and there is no data-path collective; results (a few scalars per member)
  are gathered at the end (``gather_objects``).  The reference does the same with a process pool / HTCondor map
  (ionization_scans/scan_utils.py:638-663).

* One very large SphericalHarmonicMesh simulation shards over contiguous l-blocks.  Velocity gauge: cut at EVEN channels.  The
  Crank-Nicolson solve (in r) and every even-parity l-pair sweep are local to a shard; only the odd-parity sweeps
  couple the last channel of one shard to the first channel of the next, so before each odd-parity kernel the two
  boundary channels are exchanged (3 exchanges per step; one channel = R * 16 bytes per direction) and both shards evaluate
  the straddling pair redundantly.  Length gauge: cut at ODD channels -- every odd pair (and the solve) is local, and linked
  shards run the engine's ONE-kernel step (PROG_LEN_STEP), whose read-only even-pair partners at either end of the block are the
  ghost channels: 1 exchange and one pass over psi per step.  Two transports:
  ``attach_peers`` + ``step_device`` -- the production path: the engine's own kernel stores the boundary channel into the
  neighbour's ghost channel over NVLink peer memory (CUDA IPC mapping) with a flag hand-shake, inside the captured step
  loop, no host or NCCL call per step (csrc/halo.cuh); or ``step(..., exchanger)`` -- NCCL send/recv between phases
  driven from Python (the baseline the first is measured against).
  Reductions (norm, inner products, expectation values) are per-shard partial sums + one small all-reduce.

``torch.distributed`` is plumbing only (rendezvous, NCCL p2p, all-reduce); the kernels are the engine's.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from . import engine as _engine
from . import exceptions


# ---------------------------------------------------------------------------------------------
# partitioning (pure host logic, tested on CPU)
# ---------------------------------------------------------------------------------------------
def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """contiguous block [begin, end) of ``n_items`` for ``rank`` (sizes differ by at most one)"""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def l_block_partition(l_total: int, world_size: int, cut_parity: int = 0) -> List[Tuple[int, int]]:
    """[(l_begin, L)] per rank: contiguous blocks cut at channels of one parity.

    ``cut_parity=0``: blocks begin (and, except the last, end) on EVEN channels, so that only odd-parity pairs (2m+1, 2m+2)
    straddle a cut -- every sharded program.  ``cut_parity=1`` (length gauge): blocks begin on ODD channels, every odd-parity
    pair is local to a shard, and the engine's one-pass step (even rotations folded into the odd-pair Crank-Nicolson kernel, which
    reads its even-pair partners read-only) runs on the shards with the ghost channels as those partners."""
    if l_total % 2:
        raise exceptions.UnsupportedConfiguration("l-block sharding needs an even l_bound")
    n_pairs = l_total // 2
    if world_size > n_pairs:
        raise exceptions.UnsupportedConfiguration(f"cannot cut {l_total} channels into {world_size} even blocks")
    cuts = [2 * shard_range(n_pairs, r, world_size)[0] for r in range(world_size)] + [l_total]
    if cut_parity:
        if world_size > 1 and (n_pairs - 1) < world_size:
            raise exceptions.UnsupportedConfiguration(f"cannot cut {l_total} channels into {world_size} blocks at odd channels")
        # the odd pairs (1,2) ... (l_total-3, l_total-2) are dealt out; channel 0 goes to the first block, l_total-1 to the last
        cuts = [0] + [2 * shard_range(n_pairs - 1, r, world_size)[0] + 1 for r in range(1, world_size)] + [l_total]
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world_size)]


# ---------------------------------------------------------------------------------------------
# device pointers as torch tensors (zero copy) -- for NCCL p2p on the engine's halo buffers
# ---------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr: int, n_float64: int):
        self.__cuda_array_interface__ = {"shape": (n_float64,), "typestr": "<f8", "data": (int(ptr), False), "version": 2, "strides": None}


def device_tensor(ptr: int, nbytes: int, device: int):
    import torch

    return torch.as_tensor(_CudaArray(ptr, nbytes // 8), device=torch.device("cuda", device))


# ---------------------------------------------------------------------------------------------
# halo exchange
# ---------------------------------------------------------------------------------------------
class HaloExchanger:
    """Exchanges the boundary channels of neighbouring l-block shards with torch.distributed p2p.

    ``send_lo/send_hi/recv_lo/recv_hi`` are torch tensors (views of the engine's halo buffers on CUDA; plain CPU
    tensors in the gloo tests).  ``exchange()`` posts all sends/receives of the step phase as one batch."""

    def __init__(self, rank: int, world_size: int, send_lo, send_hi, recv_lo, recv_hi, group=None):
        self.rank, self.world = rank, world_size
        self.send_lo, self.send_hi, self.recv_lo, self.recv_hi = send_lo, send_hi, recv_lo, recv_hi
        self.group = group

    def exchange(self):
        import torch.distributed as dist

        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, self.send_lo, self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_lo, self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, self.send_hi, self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_hi, self.rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


class LocalExchanger:
    """In-process stand-in used to test the shard logic on ONE GPU: all shards live in this process and the
    'exchange' is a device-to-device copy between their halo buffers."""

    def __init__(self, shards: Sequence["ShardedSimulation"]):
        self.shards = list(shards)

    def exchange_all(self):
        for lo, hi in zip(self.shards[:-1], self.shards[1:]):
            hi.recv_lo.copy_(lo.send_hi)
            lo.recv_hi.copy_(hi.send_lo)


# ---------------------------------------------------------------------------------------------
# one l-block shard of a SphericalHarmonicMesh simulation
# ---------------------------------------------------------------------------------------------
class ShardedSimulation:
    """The shard of ``problem`` (a dict of hot-path inputs, keys as in tests/golden/*.npz) owned by ``rank``."""

    def __init__(self, problem, rank: int, world_size: int, device: int = 0, use_torch_stream: bool = True, radii=(), cut_parity=None):
        kind = str(problem["kind"])
        if kind not in ("sh_len_so", "sh_vel_so"):
            raise exceptions.UnsupportedConfiguration("l-block sharding: split-operator SphericalHarmonic programs only")
        self.rank, self.world, self.device = rank, world_size, device
        L_total, R = int(problem["L"]), int(problem["R"])
        if cut_parity is None:  # length gauge: odd cuts, so that linked shards run the one-pass step (see l_block_partition)
            cut_parity = 1 if (kind == "sh_len_so" and L_total // 2 - 1 >= world_size) else 0
        self.cut_parity = int(cut_parity)
        self.l_begin, self.L = l_block_partition(L_total, world_size, self.cut_parity)[rank]
        self.L_total, self.R = L_total, R
        eng = _engine.DeviceSimulation(kind, self.L, R, batch=1, device=device, L_total=L_total, l_begin=self.l_begin)
        self.engine = eng
        if use_torch_stream:
            import torch

            self.stream = torch.cuda.Stream(device=device)
            eng.set_stream(self.stream.cuda_stream)
        else:
            self.stream = None
        lo, hi = self.l_begin - eng.g_lo, self.l_begin + self.L + eng.g_hi  # channels held, ghosts included
        eng.set_hamiltonian(np.asarray(problem["h_diag"])[lo:hi], problem["h_off"])
        if kind == "sh_vel_so":
            eng.set_vel_coupling(problem["c_l"], problem["f1_l"], problem["y_j"], problem["z_j"])
        else:
            eng.set_len_coupling(problem["c_l"], problem["x_j"])
        eng.set_mask(problem["mask"])
        sl = np.asarray(problem["state_l"]) if "state_l" in problem else np.zeros(0, dtype=np.int64)
        rows = problem["state_rows"] if len(sl) else None
        eng.set_observables(float(problem["delta_r"]), problem["r"], sl, rows, radii)
        eng.write_g(np.asarray(problem["g0"])[self.l_begin : self.l_begin + self.L].reshape(1, self.L, R))
        bufs = [eng.halo_buffer(w) for w in range(4)]
        if use_torch_stream:  # torch views of the boundary buffers: only the NCCL / in-process copy transports need them
            mk = lambda pb: None if pb[0] is None else device_tensor(pb[0], pb[1], device)
            self.send_lo, self.send_hi, self.recv_lo, self.recv_hi = (mk(b) for b in bufs)
        else:
            self._halo_bufs = bufs
        self.n_phases = eng.num_phases
        self.halo_phases = [p for p in range(self.n_phases) if eng.phase_needs_halo(p)]

    # ---- peer-memory halo exchange (the engine's own kernels over NVLink; no NCCL on the data path) --------------
    def attach_peers(self, group=None):
        """Collective: every rank exports its CUDA IPC blob, one all_gather distributes them, every rank maps its two
        neighbours.  Afterwards ``step_device`` / ``engine.run`` advance the shard with the exchange inside the engine."""
        import torch.distributed as dist

        blobs = [None] * dist.get_world_size(group)
        dist.all_gather_object(blobs, self.engine.export_peer(), group=group)
        if self.rank > 0:
            self.engine.attach_peer(0, blobs[self.rank - 1])
        if self.rank < self.world - 1:
            self.engine.attach_peer(1, blobs[self.rank + 1])
        dist.barrier(group=group)
        self.peers_attached = True

    @staticmethod
    def attach_local(shards: Sequence["ShardedSimulation"]):
        """several shards living in ONE process (tests on a single GPU): raw pointers instead of IPC handles"""
        blobs = [sh.engine.export_peer() for sh in shards]
        for i, sh in enumerate(shards):
            if i > 0:
                sh.engine.attach_peer(0, blobs[i - 1], same_process=True)
            if i + 1 < len(shards):
                sh.engine.attach_peer(1, blobs[i + 1], same_process=True)
            sh.peers_attached = True

    def step_device(self, taus, fields):
        """advance len(taus) steps with the device-resident loop (fused kernels, CUDA graphs, halo exchange by the engine's
        own kernels over peer memory).  Asynchronous; every shard must be advanced by the same number of steps."""
        self.engine.step(np.atleast_1d(taus), np.atleast_1d(fields))

    def __getattr__(self, name):
        # lazily created torch views for shards built without a torch stream (tests drive LocalExchanger with them)
        if name in ("send_lo", "send_hi", "recv_lo", "recv_hi") and "_halo_bufs" in self.__dict__:
            mk = lambda pb: None if pb[0] is None else device_tensor(pb[0], pb[1], self.device)
            vals = [mk(b) for b in self._halo_bufs]
            for k, v in zip(("send_lo", "send_hi", "recv_lo", "recv_hi"), vals):
                self.__dict__[k] = v
            return self.__dict__[name]
        raise AttributeError(name)

    def make_exchanger(self, group=None) -> HaloExchanger:
        return HaloExchanger(self.rank, self.world, self.send_lo, self.send_hi, self.recv_lo, self.recv_hi, group)

    def close(self):
        self.engine.close()

    # ---- stepping -----------------------------------------------------------------------------
    def run_phase(self, phase: int, tau: float, field: float):
        self.engine.step_phase(phase, tau, field)

    def step(self, taus, fields, exchanger: Optional[HaloExchanger] = None):
        """advance len(taus) steps; with ``exchanger`` (torch.distributed) the halos are exchanged before every
        odd-parity phase.  Every rank must call this collectively."""
        import torch

        taus = np.atleast_1d(taus)
        fields = np.atleast_1d(fields)
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _NullCtx()
        # NCCL p2p synchronises with torch's CURRENT stream.  When the engine runs on a torch stream (the default), that
        # stream is made current and everything is stream-ordered; when it runs on its own private stream the exchange is
        # fenced explicitly on both sides (the phase kernels must have finished before the boundary channels are sent, and the
        # received ghosts must have landed before the next phase reads them).
        fence = exchanger is not None and self.stream is None
        with ctx:
            for tau, f in zip(taus, fields):
                for ph in range(self.n_phases):
                    if exchanger is not None and ph in self.halo_phases:
                        if fence:
                            self.engine.synchronize()
                        exchanger.exchange()
                        if fence:
                            torch.cuda.current_stream(self.device).synchronize()
                    self.engine.step_phase(ph, float(tau), float(f))

    # ---- observables ---------------------------------------------------------------------------
    def partial_observation(self, what: int, exchanger: Optional[HaloExchanger] = None):
        """this shard's contribution to the observation record (sums over owned channels; <z> also couples the last
        owned channel to the upper ghost, which must be current: the halos are refreshed first)"""
        if what & nat.OBS_Z:
            if exchanger is not None:
                exchanger.exchange()
            elif getattr(self, "peers_attached", False):
                self.engine.exchange_halos()
        return self.engine.observe(what)[0]

    def read_g(self):
        return self.engine.read_g()[0]


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def combine_observations(records: Sequence[np.ndarray], what: int, n_states: int, l_counts: Sequence[int], n_radii: int = 0):
    """merge per-shard records into the record of the whole simulation: scalars add, norm_by_l concatenates"""
    out = []
    c = [0] * len(records)

    def take(k):
        vals = [r[c[i] : c[i] + (k[i] if isinstance(k, (list, tuple)) else k)] for i, r in enumerate(records)]
        for i in range(len(records)):
            c[i] += k[i] if isinstance(k, (list, tuple)) else k
        return vals

    if what & nat.OBS_NORM:
        out.append(np.sum(take(1), axis=0))
    if what & nat.OBS_INNER_PRODUCTS:
        out.append(np.sum(take(2 * n_states), axis=0))
    if what & nat.OBS_NORM_BY_L:
        out.append(np.concatenate(take(list(l_counts))))
    for bit in (nat.OBS_R, nat.OBS_Z, nat.OBS_H0):
        if what & bit:
            out.append(np.sum(take(1), axis=0))
    if what & nat.OBS_NORM_WITHIN:
        out.append(np.sum(take(n_radii), axis=0))
    return np.concatenate(out)


def all_reduce_observation(record: np.ndarray, device: int, group=None) -> np.ndarray:
    """sum of the additive part of an observation record over the ranks (NCCL all-reduce of a few doubles)"""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(record)).to(torch.device("cuda", device) if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def gather_objects(obj, group=None):
    """gather small python objects (per-member result dictionaries of an ensemble shard) on every rank"""
    import torch.distributed as dist

    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out
