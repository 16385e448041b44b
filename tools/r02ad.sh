#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python - <<'PY'
import os, subprocess, sys
code = r'''
import os, time, numpy as np, torch
from ionization_b200 import configs, engine, units as u
R, L, n = 16384, 4096, 40
p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=n,
                                       pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
with engine.DeviceSimulation.from_problem(p) as sim:
    st = torch.cuda.Stream(); sim.set_stream(st.cuda_stream)
    sim.step(p["taus"], p["fields"]); sim.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); sim.step(p["taus"], p["fields"]); sim.synchronize()
        ts.append(1e6 * (time.perf_counter() - t0) / n)
    print(os.environ.get("TAG"), " ".join(f"{t:.1f}" for t in ts), "us/step", flush=True)
'''
for tag, env in (("224/16", {}), ("128/16 (3 CTAs/SM)", {"ION_TSEG": "128"}), ("96/16 (4 CTAs/SM)", {"ION_TSEG": "96"}), ("352/16 (1 CTA/SM)", {"ION_TSEG": "352"})):
    e = dict(os.environ); e.update(env); e["TAG"] = tag
    subprocess.run([sys.executable, "-c", code], env=e)
PY
