"""Pulse-parameter scans: the container the reference's analysis scripts load, and a runner that shards a scan over GPUs.

* ``ParameterScan`` -- ``ionization/analysis.py:65-97``: a tag plus a list of finished simulations, stored as a gzip stream of
  pickles (first the number of simulations, then one pickle per simulation: ``ionization_scans/export_scan.py:39-47``; a
  single pickled list is accepted on load as well, ``analysis.py:84-85``).  ``parameter_set`` / ``select`` as ``analysis.py:110-123``.
  The pickles hold THIS package's classes (``ionization_b200.mesh.sims.*``); attribute names of ``sim.spec`` / ``sim.data``
  are the reference's, so scripts that go through ``ParameterScan`` and those attributes keep working.
* ``run_scan`` -- ``ionization_scans/scan_utils.py:638-663`` maps ``run(spec)`` over independent HTCondor jobs; here the specs
  are split into contiguous blocks, one per GPU (``parallel.shard_range``), every block is ONE batched device run
  (``mesh.ensemble``), and -- under ``torchrun`` -- the finished simulations (mesh stripped, as ``scan_utils.run`` returns them)
  are gathered on every rank.  No data-path collective.

    torchrun --nproc-per-node 8 -m ionization_b200.scan specs.pkl --tag my_scan [--outdir DIR]
"""
import gzip
import os
import pickle
import threading
from pathlib import Path
from typing import Any, Iterable, List, Optional, Sequence, Set

from . import exceptions, parallel


class ParameterScan:
    """analysis.py:65-123"""

    def __init__(self, tag: str, sims: Iterable):
        self.tag = tag
        self.sims = list(sims)

    @classmethod
    def from_file(cls, path, show_progress=False):
        path = Path(path).absolute()
        with gzip.open(path, mode="rb") as f:
            first = pickle.load(f)
            if isinstance(first, int):  # first entry is the number of entries
                sims = [pickle.load(f) for _ in range(first)]
            else:  # it's just a list of sims
                sims = first
        return cls(path.stem, sims)

    def save(self, outdir=None) -> Path:
        """``{outdir}/{tag}.sims`` in the format of export_scan.py:39-47 (written atomically)"""
        outdir = Path(outdir) if outdir is not None else Path.cwd()
        outpath = outdir / f"{self.tag}.sims"
        tmp = outdir / f"{self.tag}.sims.working"
        try:
            with gzip.open(tmp, mode="wb") as f:
                pickle.dump(len(self.sims), f)
                for sim in self.sims:
                    pickle.dump(sim, f)
            os.replace(tmp, outpath)
        except BaseException:
            if tmp.exists():
                tmp.unlink()
            raise
        return outpath

    def __str__(self):
        return f"{self.__class__.__name__}(tag = {self.tag})"

    def __len__(self):
        return len(self.sims)

    def __iter__(self):
        yield from self.sims

    def __getitem__(self, item):
        return self.sims[item]

    def parameter_set(self, parameter: str) -> Set[Any]:
        return {getattr(sim.spec, parameter) for sim in self.sims}

    def select(self, **parameters) -> List:
        return sorted(
            (sim for sim in self.sims if all(getattr(sim.spec, k) == v for k, v in parameters.items())),
            key=lambda sim: tuple(getattr(sim.spec, k) for k in parameters.keys()),
        )


def strip(sim):
    """what scan_utils.run returns (:656-661): the finished simulation without its mesh and without the states' numeric g"""
    for state in sim.spec.test_states:
        if hasattr(state, "g"):
            state.g = None
    sim.mesh = None
    return sim


def run_block(specs: Sequence, device: int = 0, keep_mesh: bool = False) -> List:
    """one contiguous block of a scan as ONE batched device run"""
    from .mesh import ensemble

    sims = ensemble.MeshEnsemble(specs, device=device).run() if specs else []
    return sims if keep_mesh else [strip(s) for s in sims]


def run_scan(specs: Sequence, devices: Optional[Sequence[int]] = None, group=None, keep_mesh: bool = False) -> List:
    """Run ``specs`` (simulations on one mesh and one time grid that differ in their pulse) and return the finished
    simulations in the order of ``specs``.

    * inside an initialised ``torch.distributed`` process group (one rank per GPU): rank r runs the block
      ``shard_range(len(specs), r, world)`` on ``devices[0]`` (default: LOCAL_RANK) and every rank receives all results;
    * otherwise: the blocks run concurrently on ``devices`` (default: device 0 only) from one host thread per device.
    """
    specs = list(specs)
    try:
        import torch.distributed as dist

        distributed = dist.is_available() and dist.is_initialized()
    except Exception:  # torch is optional for the single-process path
        distributed = False
    if distributed:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        b0, b1 = parallel.shard_range(len(specs), rank, world)
        device = devices[0] if devices else int(os.environ.get("LOCAL_RANK", "0"))
        mine = run_block(specs[b0:b1], device=device, keep_mesh=False)
        blocks = parallel.gather_objects((b0, mine), group=group)
        return [s for _, blk in sorted(blocks, key=lambda x: x[0]) for s in blk]
    devices = list(devices) if devices else [0]
    if len(devices) == 1:
        return run_block(specs, device=devices[0], keep_mesh=keep_mesh)
    out: List = [None] * len(devices)
    errors: List = []

    def work(i):
        b0, b1 = parallel.shard_range(len(specs), i, len(devices))
        try:
            out[i] = run_block(specs[b0:b1], device=devices[i], keep_mesh=keep_mesh)
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return [s for blk in out for s in blk]


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(description="run a pickled list of specifications as a scan sharded over the GPUs of this node")
    ap.add_argument("specs", help="pickle (optionally gzipped) of a list of specifications")
    ap.add_argument("--tag", default=None)
    ap.add_argument("--outdir", default=None)
    args = ap.parse_args(argv)
    opener = gzip.open if str(args.specs).endswith(".gz") else open
    with opener(args.specs, "rb") as f:
        specs = pickle.load(f)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            raise exceptions.NoCudaDevice("ionization_b200.scan needs one CUDA device per rank")
    sims = run_scan(specs)
    if rank == 0:
        path = ParameterScan(args.tag or Path(args.specs).stem, sims).save(args.outdir)
        print(path)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
