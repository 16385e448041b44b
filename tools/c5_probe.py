#!/usr/bin/env python
"""One-GPU run of the configs[4] mesh (or one l-block's worth of it) for ncu captures and timing: tools/c5_probe.py L_BOUND STEPS"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ionization_b200 import configs, engine  # noqa: E402
from ionization_b200 import units as u  # noqa: E402

R, L, n = 16384, int(sys.argv[1]), int(sys.argv[2])
p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=n,
                                       pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
with engine.DeviceSimulation.from_problem(p) as sim:
    st = torch.cuda.Stream()
    sim.set_stream(st.cuda_stream)
    sim.step(p["taus"], p["fields"])
    sim.synchronize()
    t0 = time.perf_counter()
    sim.step(p["taus"], p["fields"])
    sim.synchronize()
    print(f"LEN {R} x {L}: {1e6 * (time.perf_counter() - t0) / n:.1f} us/step", flush=True)
