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
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="peer: the engine's own halo kernel over NVLink peer memory inside the device-resident loop; "
                         "nccl: send/recv between phases driven from Python (the baseline)")
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
    peer = args.transport == "peer"
    ex = None if peer else shard.make_exchanger()
    if peer:
        shard.attach_peers()

    def advance(taus, fields):
        if peer:
            shard.step_device(taus, fields)
        else:
            shard.step(taus, fields, ex)

    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS
    # warm-up (NCCL communicators / graph capture, LU factors), then restore the initial state
    advance(p["taus"], p["fields"]) if peer else advance(p["taus"][:1], p["fields"][:1])
    torch.cuda.synchronize()
    shard.engine.write_g(p["g0"][shard.l_begin : shard.l_begin + shard.L].reshape(1, shard.L, R))
    dist.barrier()
    torch.cuda.synchronize()
    n_ex0 = shard.engine.halo_status()[0] if peer else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(shard.stream):
        e0.record()
    t0 = time.perf_counter()
    advance(p["taus"], p["fields"])
    with torch.cuda.stream(shard.stream):
        e1.record()
    e1.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = parallel.all_reduce_observation(shard.partial_observation(what, ex), device=local)
    g_mine = shard.read_g()

    out = {"transport": args.transport, "world": world, "r_points": R, "l_bound": L, "gauge": args.gauge, "steps": args.steps, "ms_per_step": float(ms[0]) / args.steps,
           "updates_per_s": args.steps * R * L / (float(ms[0]) * 1e-3), "norm": float(rec[0]), "halo_bytes_per_exchange_per_neighbour": shard.R * 16,
           "exchanges_per_step": (shard.engine.halo_status()[0] - n_ex0) / args.steps if peer else len(shard.halo_phases), "wall_s": wall}
    if not args.no_compare:
        gathered = parallel.gather_objects((shard.l_begin, g_mine))
        if rank == 0:
            env_keep = {k: os.environ.get(k) for k in ("ION_NO_LEN_FOLD", "ION_NO_SLAB")}
            for mode in ("same_kernels", "best"):
                # same_kernels: the unsharded run restricted to the kernel schedule the shards use (strong-scaling reference);
                # best: the unsharded run with its single-GPU-only fusions (folded LEN step, VEL slab kernel)
                if mode == "same_kernels":
                    os.environ["ION_NO_LEN_FOLD"] = "1"
                    os.environ["ION_NO_SLAB"] = "1"
                else:
                    for k, v in env_keep.items():
                        os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
                with engine.DeviceSimulation.from_problem(p, device=local) as sim:
                    sim.step(p["taus"], p["fields"])  # warm-up: factors, graph capture
                    sim.write_g(np.asarray(p["g0"], dtype=np.complex128)[None])
                    sim.synchronize()
                    t1 = time.perf_counter()
                    sim.step(p["taus"], p["fields"])
                    sim.synchronize()
                    out[f"unsharded_1gpu_ms_per_step_{mode}"] = 1e3 * (time.perf_counter() - t1) / args.steps
                    g_ref = sim.read_g()[0]
                    rec_ref = sim.observe(what)[0]
            out["speedup_vs_1gpu_same_kernels"] = out["unsharded_1gpu_ms_per_step_same_kernels"] / out["ms_per_step"]
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
