#!/usr/bin/env python
"""Parity AFTER THE FULL PULSE (BASELINE.json north_star: wavefunction, norm and ionization fraction <= 1e-10 relative): the
workload's own initial state and all of its time steps on the device (timed schedule: CUDA graphs, fused kernels) against
the oracle's C restatement on this box's host cores.  usage: tools/full_pulse_parity.py [c3_vel c3_len c1_len c4_len ...]
One JSON line per workload.  (The oracle is the checker here, as in tests/ and bench.py's parity leg.)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ionization_b200 import engine  # noqa: E402
from oracle import cport  # noqa: E402


def overlaps(problem, g):
    rows = np.asarray(problem["state_rows"])
    ls = np.asarray(problem["state_l"])
    dr = float(problem["delta_r"])
    return np.array([np.sum(np.conj(rows[k]) * g[ls[k]]) * dr for k in range(len(ls))])


for name in sys.argv[1:] or ["c3_vel", "c3_len", "c1_len", "c4_len", "c2_line"]:
    if name == "c2_line":  # configs[1]: the strongest member of the 1024-pulse scan, all 1000 time steps, 2^16 points
        wl = bench.build_workload("c2_line_ensemble")
        p = dict(wl[0])
        fields = np.ascontiguousarray(np.asarray(wl[3])[:, -1])
        t0 = time.perf_counter()
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"], fields)
            g = sim.read_g()[0, 0]
        t_gpu = time.perf_counter() - t0
        q = dict(p)
        q["fields"] = fields
        t0 = time.perf_counter()
        ref = cport.line_steps(q)
        t_cpu = time.perf_counter() - t0
        dz = float(p["delta_z"])
        norm, norm_ref = float(np.sum(np.abs(g) ** 2) * dz), float(np.sum(np.abs(ref) ** 2) * dz)
        rows = np.asarray(p["state_rows"])
        ov, ov_ref = np.abs(np.sum(np.conj(rows[0]) * g) * dz) ** 2, np.abs(np.sum(np.conj(rows[0]) * ref) * dz) ** 2
        out = {"workload": name, "time_steps": int(len(p["taus"])), "mesh": [1, int(p["Z"])], "psi_max_rel_err": float(np.max(np.abs(g - ref)) / np.max(np.abs(ref))),
               "norm": norm, "norm_rel_err": abs(norm - norm_ref) / norm_ref, "initial_state_overlap": float(ov), "ionization_fraction_rel_err": float(abs(ov - ov_ref) / max(1.0 - ov_ref, 1e-300)),
               "tolerance": 1e-10, "gpu_wall_s_incl_setup": t_gpu, "cpu_port_wall_s": t_cpu, "cpu_threads": cport.num_threads()}
        out["ok"] = bool(out["psi_max_rel_err"] <= 1e-10 and out["norm_rel_err"] <= 1e-10 and out["ionization_fraction_rel_err"] <= 1e-10)
        print(json.dumps(out), flush=True)
        continue
    wl = bench.build_workload(name)
    p = dict(wl[0])
    t0 = time.perf_counter()
    with engine.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0]
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = cport.sh_steps(p)
    t_cpu = time.perf_counter() - t0
    dr = float(p["delta_r"])
    norm, norm_ref = float(np.sum(np.abs(g) ** 2) * dr), float(np.sum(np.abs(ref) ** 2) * dr)
    bound = np.asarray(p["state_bound"], dtype=bool)
    ion, ion_ref = 1.0 - float(np.sum(np.abs(overlaps(p, g)[bound]) ** 2)), 1.0 - float(np.sum(np.abs(overlaps(p, ref)[bound]) ** 2))
    out = {"workload": name, "time_steps": int(len(p["taus"])), "mesh": [int(p["L"]), int(p["R"])],
           "psi_max_rel_err": float(np.max(np.abs(g - ref)) / np.max(np.abs(ref))),
           "norm": norm, "norm_rel_err": abs(norm - norm_ref) / norm_ref,
           "ionization_fraction_outside_test_bound_states": ion, "ionization_fraction_rel_err": abs(ion - ion_ref) / max(abs(ion_ref), 1e-300),
           "tolerance": 1e-10, "gpu_wall_s_incl_setup": t_gpu, "cpu_port_wall_s": t_cpu, "cpu_threads": cport.num_threads()}
    out["ok"] = bool(out["psi_max_rel_err"] <= 1e-10 and out["norm_rel_err"] <= 1e-10 and out["ionization_fraction_rel_err"] <= 1e-10)
    print(json.dumps(out), flush=True)
