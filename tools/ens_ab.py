#!/usr/bin/env python
"""A/B of the persistent ensemble kernel on a synthetic scan: tools/ens_ab.py R L BATCH STEPS  (run once per ION_NO_ENS value)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ionization_b200 import configs, engine, units as u

R, L, batch, steps = (int(x) for x in sys.argv[1:5])
p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=steps)
fields = np.asarray(p["fields"])[:, None] * np.linspace(0.5, 2.0, batch)[None, :]
with engine.DeviceSimulation.from_problem(p, batch=batch, with_states=False) as sim:
    sim.step(p["taus"], fields); sim.synchronize()
    t0 = time.perf_counter()
    sim.step(p["taus"], fields); sim.synchronize()
    dt = time.perf_counter() - t0
print(f"ION_NO_ENS={os.environ.get('ION_NO_ENS', '0')} R={R} L={L} batch={batch}: {1e6 * dt / steps:.1f} us/step, {steps * R * L * batch / dt / 1e9:.1f} G updates/s")
