import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from ionization_b200 import configs, engine, parallel
from ionization_b200 import units as u
import test_gpu_shards_and_segments as T
R, L, world, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(os.environ.get("STEPS", "12"))
p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge=sys.argv[4] if len(sys.argv) > 4 else "LEN", n_steps=n,
                                       pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
rng = np.random.default_rng(0)
g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2) * 3)[None, :]
p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))
p["state_l"] = np.zeros(0, dtype=np.int64)
with engine.DeviceSimulation.from_problem(p, with_states=False) as sim:
    sim.step(p["taus"], p["fields"])
    ref = sim.read_g()[0]
def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
from oracle import cport
orc = cport.sh_steps(p)
print("unsharded vs oracle", rel(ref, orc), flush=True)
with engine.DeviceSimulation.from_problem(p, with_states=False) as sim:
    sim.step(p["taus"], p["fields"])
    sim.write_g(np.asarray(p["g0"], dtype=np.complex128)[None])
    sim.step(p["taus"], p["fields"])
    print("unsharded second run vs oracle", rel(sim.read_g()[0], orc), flush=True)
for name, fn in (("local/phases", T._run_sharded), ("peer/device", T._run_sharded_device)):
    g, _ = fn(p, world)
    d = np.abs(g - ref)
    l_bad, r_bad = np.unravel_index(np.argmax(d), d.shape)
    print(name, "rel err vs unsharded", rel(g, ref), "worst at l, r =", l_bad, r_bad, flush=True)
