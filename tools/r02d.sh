#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import time, os, numpy as np
from ionization_b200 import configs, engine
nat = engine.nat
for gauge in ("VEL", "LEN"):
    p = configs.config3(gauge)
    n = 512
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_NORM_BY_L
    for mode, env, mask in (("unobserved", "0", np.zeros(n, np.uint8)), ("every step, fused", "0", np.ones(n, np.uint8)), ("every step, separate", "1", np.ones(n, np.uint8))):
        os.environ["ION_NO_FUSED_OBS"] = env
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.run(p["taus"][:n], p["fields"][:n], mask, what)
            sim.synchronize()
            t0 = time.perf_counter()
            sim.run(p["taus"][:n], p["fields"][:n], mask, what)
            sim.synchronize()
            print(f"C3 {gauge} {mode}: {1e6 * (time.perf_counter() - t0) / n:.2f} us/step", flush=True)
PY
