"""Host-side logic of the multi-GPU modes, on the CPU: partitioning, and the halo-exchange schedule between two
ranks over the gloo backend (world_size 2)."""
import os
import socket

import numpy as np
import pytest

from ionization_b200 import parallel
from ionization_b200 import _native as nat


def test_shard_range_covers_everything_once():
    for n in (1, 7, 64, 4096):
        for w in (1, 2, 3, 8):
            blocks = [parallel.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_l_block_partition_cuts_at_even_channels():
    for L in (10, 200, 500, 4096):
        for w in (1, 2, 4, 5):
            blocks = parallel.l_block_partition(L, w)
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == L
            for (b0, n0), (b1, _) in zip(blocks[:-1], blocks[1:]):
                assert b0 + n0 == b1 and b1 % 2 == 0
    with pytest.raises(Exception):
        parallel.l_block_partition(11, 2)
    with pytest.raises(Exception):
        parallel.l_block_partition(4, 3)


def test_l_block_partition_cuts_at_odd_channels():
    """length gauge: every odd pair (2m+1, 2m+2) stays inside one block; channel 0 and the last channel go to the end blocks"""
    for L in (10, 200, 500, 4096):
        for w in (1, 2, 4, 5):
            if L // 2 - 1 < w and w > 1:
                continue
            blocks = parallel.l_block_partition(L, w, cut_parity=1)
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == L
            for (b0, n0), (b1, _) in zip(blocks[:-1], blocks[1:]):
                assert b0 + n0 == b1 and b1 % 2 == 1
            sizes = [n for _, n in blocks]
            assert max(sizes) - min(sizes) <= 3
    with pytest.raises(Exception):
        parallel.l_block_partition(4, 2, cut_parity=1)


def test_combine_observations():
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_NORM_BY_L | nat.OBS_R
    a = np.array([0.25, 1.0, 2.0, 0.0, 0.0, 0.1, 0.15, 3.0])  # norm, ip0(re,im), ip1, nbl x2, r
    b = np.array([0.5, 0.0, 0.0, 0.5, -0.5, 0.2, 0.3, 4.0])
    out = parallel.combine_observations([a, b], what, n_states=2, l_counts=[2, 2])
    assert np.allclose(out, [0.75, 1.0, 2.0, 0.5, -0.5, 0.1, 0.15, 0.2, 0.3, 7.0])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_worker(rank, world, port, results):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 16
        send_lo = torch.full((n,), 10.0 * rank + 1, dtype=torch.float64)  # my first channel
        send_hi = torch.full((n,), 10.0 * rank + 2, dtype=torch.float64)  # my last channel
        recv_lo = torch.full((n,), -1.0, dtype=torch.float64)
        recv_hi = torch.full((n,), -1.0, dtype=torch.float64)
        ex = parallel.HaloExchanger(rank, world, send_lo, send_hi, recv_lo, recv_hi)
        for step in range(3):
            send_lo += 100
            send_hi += 100
            ex.exchange()
        # partial observation records add up
        rec = parallel.all_reduce_observation(np.array([rank + 1.0, 2.0 * rank]), device=0)
        results[rank] = (recv_lo[0].item(), recv_hi[0].item(), rec.tolist())
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world_size_2_and_3_gloo():
    import torch.multiprocessing as mp

    for world in (2, 3):
        port = _free_port()
        mgr = mp.Manager()
        results = mgr.dict()
        mp.spawn(_halo_worker, args=(world, port, results), nprocs=world, join=True)
        for rank in range(world):
            lo, hi, rec = results[rank]
            # lower ghost = last channel of rank-1 after 3 steps; upper ghost = first channel of rank+1
            assert lo == (10.0 * (rank - 1) + 2 + 300 if rank > 0 else -1.0)
            assert hi == (10.0 * (rank + 1) + 1 + 300 if rank < world - 1 else -1.0)
            assert rec == [sum(r + 1.0 for r in range(world)), sum(2.0 * r for r in range(world))]


class _Spec:
    def __init__(self, k):
        self.k = k
        self.name = f"m{k}"


class _Sim:
    def __init__(self, spec, device):
        self.spec, self.device, self.mesh = spec, device, None


def _fake_block(specs, device=0, keep_mesh=False):
    return [_Sim(s, device) for s in specs]


def _scan_worker(rank, world, port, results):
    import torch.distributed as dist

    from ionization_b200 import scan

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scan.run_block = _fake_block  # the batched device run of a block is GPU work; the sharding and the gather are what is tested here
        sims = scan.run_scan([_Spec(k) for k in range(7)])
        results[rank] = [(s.spec.k, s.device) for s in sims]
    finally:
        dist.destroy_process_group()


def test_scan_is_split_into_contiguous_blocks_and_gathered_on_every_rank_gloo():
    """ionization_b200.scan.run_scan under a process group (world_size 2 and 3): rank r runs block shard_range(n, r, world) on its
    LOCAL_RANK device and every rank receives all simulations in the order of the specs (scan_utils.py:638-663 maps run(spec))"""
    import torch.multiprocessing as mp

    for world in (2, 3):
        port = _free_port()
        mgr = mp.Manager()
        results = mgr.dict()
        mp.spawn(_scan_worker, args=(world, port, results), nprocs=world, join=True)
        expect = []
        for r in range(world):
            b0, b1 = parallel.shard_range(7, r, world)
            expect += [(k, r) for k in range(b0, b1)]
        for rank in range(world):
            assert results[rank] == expect


def test_scan_over_several_devices_of_one_process_keeps_the_order(monkeypatch):
    from ionization_b200 import scan

    monkeypatch.setattr(scan, "run_block", _fake_block)
    sims = scan.run_scan([_Spec(k) for k in range(10)], devices=[3, 5, 6])
    assert [s.spec.k for s in sims] == list(range(10))
    assert [s.device for s in sims] == [3] * 4 + [5] * 3 + [6] * 3
