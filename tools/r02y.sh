#!/bin/bash
# A/B: HEAD build vs the working tree (recomputed halos, cluster off) on C5 one GPU; C3 sanity
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
for tag, env in (("head", {"ION_LIB": "/root/repo/ionization_b200/_lib/exp_head.so"}), ("tree halo", {"ION_NO_CLUSTER": "1"}),
                 ("head", {"ION_LIB": "/root/repo/ionization_b200/_lib/exp_head.so"}), ("tree halo", {"ION_NO_CLUSTER": "1"}), ("tree cluster", {})):
    e = dict(os.environ); e.update(env); e["TAG"] = tag
    subprocess.run([sys.executable, "-c", code], env=e)
PY
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
tools/ab_env.sh c3_vel 1000 "ION_LIB=/root/repo/ionization_b200/_lib/exp_head.so" "X=1"
tools/ab_env.sh c3_len 1000 "ION_LIB=/root/repo/ionization_b200/_lib/exp_head.so" "X=1"
