#!/usr/bin/env python
"""Time the on-chip resident kernel with parts switched off (ION_RES_DBG bits: 1 no waiting for the neighbours, 2 no
Crank-Nicolson, 4 no l-sweeps) to attribute the step time.  Results with any bit set are wrong by construction.
usage: tools/res_probe.py [workload ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ionization_b200 import configs, engine  # noqa: E402

n_t = int(os.environ.get("PROBE_STEPS", "400"))
for wl in sys.argv[1:] or ["c3_len", "c3_vel"]:
    p = configs.config3("VEL" if wl.endswith("vel") else "LEN") if wl.startswith("c3") else configs.config1("VEL" if wl.endswith("vel") else "LEN")
    taus, fields = np.ascontiguousarray(p["taus"][:n_t]), np.ascontiguousarray(p["fields"][:n_t])
    sim = engine.DeviceSimulation.from_problem(p)
    g0 = np.asarray(p["g0"], dtype=np.complex128)[None]
    for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
        os.environ["ION_RES_DBG"] = str(dbg)
        best = 1e9
        for _ in range(3):
            sim.write_g(g0)
            sim.synchronize()
            t0 = time.perf_counter()
            sim.step(taus, fields)
            sim.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"{wl} dbg={dbg} ({'nowait ' if dbg & 1 else ''}{'nocn ' if dbg & 2 else ''}{'nosweep' if dbg & 4 else ''}): {1e6 * best / n_t:.2f} us/step", flush=True)
    os.environ["ION_RES_DBG"] = "0"
    sim.close()
