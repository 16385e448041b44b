import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from ionization_b200 import configs, engine, parallel
from ionization_b200 import units as u
from oracle import cport
R, L, n = int(sys.argv[1]), int(sys.argv[2]), int(os.environ.get("STEPS", "12"))
p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=n,
                                       pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
rng = np.random.default_rng(0)
g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2) * 3)[None, :]
p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))
ref = cport.sh_steps(p)
def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
for env in ({}, {"ION_NO_LEN_FOLD": "1"}, {"ION_NO_LEN_FOLD": "1", "ION_NO_GRAPHS": "1"}):
    os.environ.update(env)
    with engine.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0]
    for k in env: del os.environ[k]
    d = np.abs(g - ref)
    l_bad, r_bad = np.unravel_index(np.argmax(d), d.shape)
    print(env, "rel err vs oracle", rel(g, ref), "worst at l, r =", l_bad, r_bad, flush=True)
