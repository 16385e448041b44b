"""One SphericalHarmonicMesh simulation on several GPUs of this process, behind the mesh API:

    SphericalHarmonicSpecification(..., devices=[0, 1, 2, 3]).to_sim().run()

The wavefunction is cut into contiguous l-blocks at even channels (``parallel.l_block_partition``), one ``DeviceSimulation`` per
GPU; the shards are linked once with the engine's peer-memory halo exchange (``csrc/halo.cuh``: boundary channels are stored into
the neighbour's memory over NVLink by the engine's own kernel inside the captured step loop) and advanced from one host thread
each -- the same rendezvous a rank per GPU performs under ``torchrun`` (``parallel.ShardedSimulation.attach_peers``).
``ShardedEngine`` has the interface ``MeshSimulation`` / ``QuantumMesh`` use of ``engine.DeviceSimulation``; observation records
are per-shard partial sums combined on the host (``parallel.combine_observations``).  Split-operator programs, even ``l_bound``;
<z> and the energies couple a shard's last channel to its neighbour's current first channel and are not available here.
"""
import threading

import numpy as np

from .. import _native as nat
from .. import exceptions, parallel

_UNSUPPORTED = nat.OBS_Z | nat.OBS_H0


class ShardedEngine:
    def __init__(self, problem, devices, radii=()):
        self.devices = [int(d) for d in devices]
        world = len(self.devices)
        self.L, self.R, self.batch = int(problem["L"]), int(problem["R"]), 1
        self.n_states = len(problem["state_l"]) if "state_l" in problem else 0
        self.n_radii = len(radii)
        self.shards = []
        try:
            for rank, dev in enumerate(self.devices):
                self.shards.append(parallel.ShardedSimulation(problem, rank, world, device=dev, use_torch_stream=False, radii=radii))
            parallel.ShardedSimulation.attach_local(self.shards)
        except BaseException:
            self.close()
            raise
        self._prepared_tau = None

    # -- plumbing ---------------------------------------------------------------------------------
    def _each(self, fn):
        """fn(shard) on every shard from its own host thread (calls that enqueue halo exchanges must be concurrent: the exchange
        is a rendezvous on the device); returns the results in shard order"""
        out, errors = [None] * len(self.shards), []

        def work(i, sh):
            try:
                out[i] = fn(sh)
            except BaseException as exc:  # noqa: BLE001
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(i, sh)) for i, sh in enumerate(self.shards)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out

    def _prepare(self, taus, n_records=0, what=0):
        """everything that allocates or frees happens here, shard by shard, before the concurrent part: cudaFree waits for the
        device to go idle, which it never does while a neighbour's exchange kernel is waiting for this shard's kernels"""
        taus = np.atleast_1d(taus)
        tau = float(taus[0])
        for sh in self.shards:
            sh.engine.synchronize()
        for sh in self.shards:
            if self._prepared_tau != tau:
                sh.engine.prepare(tau)
            sh.engine.reserve(len(taus), n_records, what)
        self._prepared_tau = tau

    def _check(self, what):
        if what & _UNSUPPORTED:
            raise exceptions.UnsupportedConfiguration("<z> and energy expectation values are not available for an l-block sharded simulation (devices=[...])")

    def _combine(self, recs, what):
        return parallel.combine_observations(recs, what, n_states=self.n_states, l_counts=[sh.L for sh in self.shards], n_radii=self.n_radii)

    def close(self):
        for sh in self.shards:
            try:
                sh.close()
            except Exception:  # noqa: BLE001
                pass
        self.shards = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the DeviceSimulation interface used by the mesh layer ----------------------------------------
    def set_mask(self, mask):
        for sh in self.shards:
            sh.engine.set_mask(mask)

    def write_g(self, g):
        g = np.asarray(g, dtype=np.complex128).reshape(self.L, self.R)
        for sh in self.shards:
            sh.engine.write_g(np.ascontiguousarray(g[sh.l_begin : sh.l_begin + sh.L]).reshape(1, sh.L, self.R))

    def read_g(self, out=None):
        g = np.concatenate([sh.read_g() for sh in self.shards], axis=0).reshape(1, self.L, self.R)
        if out is not None:
            out[...] = g
            return out
        return g

    def step(self, taus, fields):
        self._prepare(taus)
        self._each(lambda sh: sh.engine.step(np.atleast_1d(taus), np.atleast_1d(fields)))

    def run(self, taus, fields, observe_mask=None, what=0):
        self._check(what)
        self._prepare(taus, 0 if observe_mask is None else int(np.count_nonzero(observe_mask)), what)
        recs = self._each(lambda sh: sh.engine.run(np.atleast_1d(taus), np.atleast_1d(fields), observe_mask, what))
        n_obs = recs[0].shape[0]
        size = self.observation_size(what)
        out = np.empty((n_obs, 1, size), dtype=np.float64)
        for k in range(n_obs):
            out[k, 0] = self._combine([r[k, 0] for r in recs], what)
        return out

    def observation_size(self, what):
        n = 0
        n += 1 if what & nat.OBS_NORM else 0
        n += 2 * self.n_states if what & nat.OBS_INNER_PRODUCTS else 0
        n += self.L if what & nat.OBS_NORM_BY_L else 0
        n += 1 if what & nat.OBS_R else 0
        n += self.n_radii if what & nat.OBS_NORM_WITHIN else 0
        return n

    def observe(self, what):
        self._check(what)
        recs = [sh.engine.observe(what)[0] for sh in self.shards]
        return self._combine(recs, what).reshape(1, -1)

    def synchronize(self):
        for sh in self.shards:
            sh.engine.synchronize()

    @property
    def launch_count(self):
        return sum(sh.engine.launch_count for sh in self.shards)
