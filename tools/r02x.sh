#!/bin/bash
# clustered r-segments: how does the cost of a cluster scale with its size?  (cluster vs recomputed halos, HBM-resident meshes)
timeout 800 python - <<'PY'
import os, time, numpy as np, torch
from ionization_b200 import configs, engine, units as u
os.environ["ION_DEBUG"] = "1"
for gauge, R, L, n in (("LEN", 5120, 8192, 40), ("LEN", 8192, 8192, 40), ("LEN", 12288, 4096, 40), ("LEN", 16384, 4096, 40)):
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge=gauge, n_steps=n,
                                           pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
    for env in ({"ION_NO_CLUSTER": "1"}, {}, {"ION_NO_PDL": "1"}, {"ION_NO_CLUSTER": "1"}):
        for k in ("ION_NO_CLUSTER", "ION_CLUSTER_TSEG", "ION_NO_PDL"): os.environ.pop(k, None)
        os.environ.update(env)
        try:
            with engine.DeviceSimulation.from_problem(p) as sim:
                st = torch.cuda.Stream(); sim.set_stream(st.cuda_stream)
                sim.step(p["taus"], p["fields"]); sim.synchronize()
                t0 = time.perf_counter(); sim.step(p["taus"], p["fields"]); sim.synchronize()
                dt = 1e6 * (time.perf_counter() - t0) / n
                print(gauge, R, L, env, f"{dt:.1f} us/step  {R * L / dt * 1e-3:.1f} G/s", flush=True)
        except Exception as e:
            print(gauge, R, L, env, "FAILED", repr(e)[:300], flush=True)
PY
