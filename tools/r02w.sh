#!/bin/bash
# clustered r-segments (DSMEM hand-off of the scan inflow): parity against the oracle, then timing of C5 on one GPU
for shape in "5000 6" "8192 8" "16384 6"; do
  timeout 300 python tools/seg_check.py $shape 2>&1 | tail -4
done
timeout 600 python - <<'PY'
import os, time, numpy as np, torch
from ionization_b200 import configs, engine, units as u
for gauge, R, L, n in (("LEN", 16384, 4096, 40),):
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge=gauge, n_steps=n,
                                           pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
    for env in ({}, {"ION_NO_CLUSTER": "1"}, {"ION_CLUSTER_TSEG": "512"}, {"ION_CLUSTER_TSEG": "384"}):
        for k in ("ION_NO_CLUSTER", "ION_CLUSTER_TSEG"): os.environ.pop(k, None)
        os.environ.update(env)
        try:
            with engine.DeviceSimulation.from_problem(p) as sim:
                st = torch.cuda.Stream(); sim.set_stream(st.cuda_stream)
                sim.step(p["taus"], p["fields"]); sim.synchronize()
                t0 = time.perf_counter(); sim.step(p["taus"], p["fields"]); sim.synchronize()
                print(gauge, R, L, env, f"{1e6 * (time.perf_counter() - t0) / n:.1f} us/step", flush=True)
        except Exception as e:
            print(gauge, R, L, env, "FAILED", repr(e)[:300], flush=True)
PY
