#!/bin/bash
# persistent (pair, segment) pipeline for one long length-gauge simulation (ensemble.cuh SEGM) vs one CTA per task: parity, C5 timing
for shape in "5000 400" "9000 300"; do
  timeout 300 python tools/seg_check.py $shape 2>&1 | head -1
done
timeout 300 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -x -q -k "c5 or C5 or 16384" 2>&1 | tail -2
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
for tag, env in (("persistent 224/16", {}), ("one CTA per task 224/16", {"ION_NO_SEG_ENS": "1"}), ("persistent 224/16", {})):
    e = dict(os.environ); e.update(env); e["TAG"] = tag
    subprocess.run([sys.executable, "-c", code], env=e)
PY
