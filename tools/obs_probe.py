#!/usr/bin/env python
"""every-step observation on the C3 mesh: us per time step on the fused schedule vs the single-sweep schedule (ION_NO_FUSED_OBS=1).
usage: tools/obs_probe.py [VEL|LEN] [n_steps]   (run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionization_b200 import configs, engine  # noqa: E402

gauge = sys.argv[1] if len(sys.argv) > 1 else "VEL"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
nat = engine.nat
p = configs.config3(gauge)
what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_NORM_BY_L
for mode, mask in (("unobserved", np.zeros(n, np.uint8)), ("every step", np.ones(n, np.uint8))):
    with engine.DeviceSimulation.from_problem(p) as sim:
        sim.run(p["taus"][:n], p["fields"][:n], mask, what)
        sim.synchronize()
        t0 = time.perf_counter()
        sim.run(p["taus"][:n], p["fields"][:n], mask, what)
        sim.synchronize()
        print(f"C3 {gauge} {mode} (ION_NO_FUSED_OBS={os.environ.get('ION_NO_FUSED_OBS', '0')}): {1e6 * (time.perf_counter() - t0) / n:.2f} us/step", flush=True)
