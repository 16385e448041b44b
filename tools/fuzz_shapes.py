#!/usr/bin/env python
"""Random mesh sizes through the kernel variants that depend on the geometry -- TMA-staged pair kernels (288..512-thread CTAs),
half-warp-halo r-segments, the eight-row LineMesh Crank-Nicolson kernel, the eight-row ADI radial solve -- against the oracle.
usage: tools/fuzz_shapes.py [SEED] [CASES_PER_KIND]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ionization_b200 import configs, engine  # noqa: E402
from ionization_b200 import units as u  # noqa: E402
from oracle import cport, restate  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(seed)
worst = 0.0


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def sh_case(gauge, R, L, n, adi=False):
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge=gauge, n_steps=n,
                                           pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
    g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2) * 3)[None, :]
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))
    if adi:
        p["kind"] = "sh_len_adi"
        ref = restate.run_sh(p, store_every_step=False)["g"]
    else:
        ref = cport.sh_steps(p)
    with engine.DeviceSimulation.from_problem(p, with_states=False) as sim:
        sim.step(p["taus"], p["fields"])
        return rel(sim.read_g()[0], ref)


def line_case(Z, n):
    from conftest import load_golden

    base = dict(load_golden("line_len_cn_1024"))
    z = np.linspace(-1, 1, Z) * base["z"][-1] * (Z / 1024)
    dz = z[1] - z[0]
    scale = (float(base["delta_z"]) / dz) ** 2
    p = dict(base)
    p.update(Z=Z, z=z, delta_z=dz, h_off=np.full(Z - 1, base["h_off"][0] * scale), w_z=z * (base["w_z"][-1] / base["z"][-1]), mask=np.cos(np.linspace(0, 1.0, Z)) ** 0.125)
    p["h_diag"] = np.full(Z, -2 * p["h_off"][0]) + 0j + np.interp(z, base["z"], np.real(base["h_diag"]) + 2 * base["h_off"][0])
    g0 = (rng.standard_normal(Z) + 1j * rng.standard_normal(Z)) * np.exp(-((z / z[-1]) ** 2) * 2)
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * dz)
    p["state_rows"] = p["g0"][None, :]
    p["taus"], p["fields"] = p["taus"][:n], 0.05 * np.asarray(p["fields"][:n])
    ref = cport.line_steps(p)
    with engine.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        return rel(sim.read_g()[0, 0], ref)


kinds = [
    ("pair kernels, TMA-staged LU factors (LEN)", lambda: ("LEN", int(rng.integers(1153, 2049)), 2 * int(rng.integers(2, 9)), int(rng.integers(3, 9)))),
    ("pair kernels, TMA-staged LU factors (VEL)", lambda: ("VEL", int(rng.integers(1153, 2049)), 2 * int(rng.integers(2, 9)), int(rng.integers(3, 9)))),
    ("r-segments with half-warp halos (LEN)", lambda: ("LEN", int(rng.integers(4097, 12000)), 2 * int(rng.integers(2, 6)), int(rng.integers(3, 7)))),
    ("r-segments (VEL)", lambda: ("VEL", int(rng.integers(4097, 9000)), 2 * int(rng.integers(2, 5)), int(rng.integers(3, 6)))),
]
for name, draw in kinds:
    for _ in range(n_cases):
        gauge, R, L, n = draw()
        e = sh_case(gauge, R, L, n)
        worst = max(worst, e)
        print(f"{name}: R={R} L={L} steps={n}  rel err {e:.2e}", flush=True)
for _ in range(n_cases):
    R, L, n = int(rng.integers(130, 2049)), int(rng.integers(3, 40)), int(rng.integers(3, 8))
    e = sh_case("LEN", R, L, n, adi=True)
    worst = max(worst, e)
    print(f"ADI (k_adi_r for T <= 512): R={R} L={L} steps={n}  rel err {e:.2e}", flush=True)
for _ in range(n_cases):
    Z, n = int(rng.integers(4097, 30000)), int(rng.integers(3, 8))
    e = line_case(Z, n)
    worst = max(worst, e)
    print(f"LineMesh CN, eight rows per thread: Z={Z} steps={n}  rel err {e:.2e}", flush=True)
print(f"worst relative error over all cases: {worst:.2e}  ({'ok' if worst < 1e-10 else 'FAILED'}: tolerance 1e-10)")
sys.exit(0 if worst < 1e-10 else 1)
