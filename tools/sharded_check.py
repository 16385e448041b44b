#!/usr/bin/env python
"""Multi-GPU check of l-block sharding with NCCL halo exchange (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_check.py [--r-points 4096] [--l-bound 512] [--steps 20] [--gauge LEN]

Every rank evolves its l-block; rank 0 also evolves the whole mesh unsharded on its GPU and compares
(parity of the sharded path) and prints timing of the sharded stepping loop (device time, max over ranks).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--r-points", type=int, default=4096)
    ap.add_argument("--l-bound", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--gauge", default="LEN")
    ap.add_argument("--no-compare", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from ionization_b200 import _native as nat
    from ionization_b200 import configs, engine, parallel
    from ionization_b200 import units as u

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    R, L = args.r_points, args.l_bound
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge=args.gauge, n_steps=args.steps,
                                           pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-args.steps / 2 * u.asec,
                                           time_final=args.steps / 2 * u.asec)
    # populate many channels so the cuts carry amplitude
    rng = np.random.default_rng(0)
    g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2) * 3)[None, :]
    g0 *= np.exp(-np.arange(L) / (L / 4))[:, None]
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))

    shard = parallel.ShardedSimulation(p, rank, world, device=local)
    ex = shard.make_exchanger()
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS
    # warm-up step (NCCL communicators, LU factors), then restore the initial state
    shard.step(p["taus"][:1], p["fields"][:1], ex)
    torch.cuda.synchronize()
    shard.engine.write_g(p["g0"][shard.l_begin : shard.l_begin + shard.L].reshape(1, shard.L, R))
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(shard.stream):
        e0.record()
    t0 = time.perf_counter()
    shard.step(p["taus"], p["fields"], ex)
    with torch.cuda.stream(shard.stream):
        e1.record()
    e1.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = parallel.all_reduce_observation(shard.partial_observation(what, ex), device=local)
    g_mine = shard.read_g()

    out = {"world": world, "r_points": R, "l_bound": L, "gauge": args.gauge, "steps": args.steps, "ms_per_step": float(ms[0]) / args.steps,
           "updates_per_s": args.steps * R * L / (float(ms[0]) * 1e-3), "norm": float(rec[0]), "halo_bytes_per_exchange_per_neighbour": shard.R * 16,
           "exchanges_per_step": len(shard.halo_phases), "wall_s": wall}
    if not args.no_compare:
        gathered = parallel.gather_objects((shard.l_begin, g_mine))
        if rank == 0:
            with engine.DeviceSimulation.from_problem(p, device=local) as sim:
                sim.step(p["taus"], p["fields"])
                g_ref = sim.read_g()[0]
                rec_ref = sim.observe(what)[0]
            g = np.concatenate([blk for _, blk in sorted(gathered, key=lambda x: x[0])], axis=0)
            out["max_rel_err_vs_unsharded"] = float(np.max(np.abs(g - g_ref)) / np.max(np.abs(g_ref)))
            out["norm_err"] = abs(float(rec[0]) - float(rec_ref[0]))
            out["ip_err"] = float(np.max(np.abs(rec[1:] - rec_ref[1:])))
    if rank == 0:
        print(json.dumps(out), flush=True)
        if "max_rel_err_vs_unsharded" in out:
            assert out["max_rel_err_vs_unsharded"] < 1e-10 and out["norm_err"] < 1e-10 and out["ip_err"] < 1e-10, out
    shard.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
