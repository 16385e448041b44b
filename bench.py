#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the mesh time-evolution hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3_vel|c3_len|c3_adi|c1_len|c4_len|c4_len_ensemble|c2_line_ensemble]

Metric (BASELINE.json): grid-point updates/s of the SphericalHarmonicMesh CN + split-operator step, and the
fraction of the B200 HBM roofline at 32 B per update.

A bench "step" is ONE PASS OF THE HOT PATH OVER ONE BATCH OF INPUT = one MeshSimulation.run() of the workload's
pulse: ``time_steps`` consecutive time steps (2000 for configs[2]) of the whole mesh.  So
    value = K * time_steps * mesh_points * sims_per_gpu * N / seconds(K steps)       [grid-point updates / s]
  * ``value``: inputs resident in HBM when the timed region starts (psi reset on the device outside the timed region);
  * ``e2e``:  the same through the public host-buffer API -- every step copies psi_0 and the per-step field scalars
    host->device (pinned memory), evolves, and reads psi_final and an observation record (norm + inner products)
    device->host, all inside the timed region.
N > 1 (torchrun, one rank per GPU): scan-ensemble sharding -- every rank evolves its own independent simulation(s),
no data-path collective; time = max over ranks; "scaling": "weak".

--impl reference: the reference's CPU path for the same workload, timed on this box's host cores.  The reference is a
Python package that cannot travel to the GPU box (it needs the un-vendored `simulacra`), so its path is represented by
the oracle's C restatement (oracle/c/restate.c, OpenMP over all host threads; "kind": "port"), pinned to the real
reference by tests/golden.  Each reference step is a bounded sample (a few time steps of the same mesh).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UPDATE = 32.0  # one complex128 read + one write of psi per time step


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture of this
    workload (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when no capture is on file."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            entry = json.load(f)[workload][kernel]
        return float(entry["dram_bytes_read"]) + float(entry["dram_bytes_write"])
    except Exception:
        return None


def ncu_traffic_warm(workload, kernel):
    """the same from the capture without cache flushes between replays (the L2-resident steady state), if on file"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            entry = json.load(f)[workload][kernel]
        return float(entry["warm_dram_bytes_read"]) + float(entry["warm_dram_bytes_write"])
    except Exception:
        return None


def fp64_inst_per_point(workload, prof, n_prof, points):
    """FP64 instructions (thread level: DFMA + DMUL + DADD) per grid point and time step: sum over the kernels of a step of
    (instructions per launch from the committed ncu capture) x (launches per time step measured now) / points"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            entry = json.load(f)[workload]
    except Exception:
        entry = {}
    total, used = 0.0, []
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    for k, (ms_k, n_k) in prof.items():
        kname = {"slab": "k_slab", "adi_l": "k_adi_l", "len_ens": "k_len_ens"}.get(k, f"k_unit<{k}>")
        e = entry.get(kname, {})
        if "fp64_thread_inst" in e:
            total += float(e["fp64_thread_inst"]) * (n_k / n_prof)
            used.append(kname)
        elif ms_k / total_ms > 0.05:  # a kernel with a real share of the step and no capture: fall back to the static count
            return STATIC_FP64_PER_POINT.get(workload), "static count (DESIGN.md section 6); no ncu capture on file for " + kname
    return (total / points if used else STATIC_FP64_PER_POINT.get(workload)), ("ncu capture (profiles/ncu_traffic.json): " + ", ".join(used)) if used else "static count (DESIGN.md section 6)"


STATIC_FP64_PER_POINT = {"c3_vel": 153.0, "c3_len": 66.0, "c4_len_ensemble": 66.0}  # rounded from the ncu captures of round 2 (profiles/ncu_traffic.json)


def build_workload(name):
    """-> (problem, description[, batch, fields[n_steps, batch]])"""
    from ionization_b200 import configs

    if name == "c4_len_ensemble":
        # configs[3] per GPU: 512 members of the 64 x 64 fluence x CEP scan on r_points=1000, l_bound=200
        p = configs.config4_member("LEN")
        fields = configs.scan_fields(p, np.geomspace(0.01, 20, 16), np.linspace(0, 2 * np.pi, 32, endpoint=False))
        return p, "configs[3] (one GPU's share): 512-member fluence x CEP scan, SphericalHarmonicMesh r_points=1000 l_bound=200 length gauge split-operator, 2000 steps", 512, fields
    if name == "c2_line_ensemble":
        p = configs.config2(n_steps=1000)
        fields = configs.scan_fields(p, np.geomspace(0.01, 10, 32), np.linspace(0, 2 * np.pi, 32, endpoint=False))
        return p, "configs[1]: LineMesh Gaussian well 2^16 points Crank-Nicolson length gauge, batch of 1024 Sinc pulses, 1000 steps", 1024, fields
    if name == "c3_vel":
        return configs.config3("VEL"), "configs[2]: SphericalHarmonicMesh hydrogen 1s r_bound=250a0 r_points=2000 l_bound=500 velocity gauge split-operator, Sinc 200as, 2000 steps, single sim"
    if name == "c3_len":
        return configs.config3("LEN"), "SphericalHarmonicMesh r_points=2000 l_bound=500 length gauge split-operator, Sinc 200as, 2000 steps, single sim"
    if name == "c3_adi":
        p = dict(configs.config3("LEN"))
        p["kind"] = "sh_len_adi"  # same mesh, pulse and field samples (E(t + dt/2), mesh_operators.py:1011-1013); evolution_methods.py:46-77
        return p, "SphericalHarmonicMesh r_points=2000 l_bound=500 length gauge AlternatingDirectionImplicit, Sinc 200as, 2000 steps, single sim"
    if name == "c1_len":
        return configs.config1("LEN"), "configs[0]: SphericalHarmonicMesh r_points=500 l_bound=50 length gauge split-operator, 2000 steps"
    if name == "c4_len":
        return configs.config4_member("LEN"), "configs[3] member: SphericalHarmonicMesh r_points=1000 l_bound=200 length gauge split-operator"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


PORT_NOTE = ("the reference package itself (numpy / scipy.sparse / one Cython tdma, single core; it cannot travel to the GPU box) measured 0.93 M updates/s on "
             "configs[2] and 2.3 M updates/s on configs[0] in the build container (BASELINE.md section 2): this C/OpenMP restatement is ~76x faster than the package")


def cpu_steps(problem, n, members=1):
    """n time steps of the workload on the CPU port; returns grid-point updates done"""
    from oracle import cport

    if str(problem["kind"]).startswith("line"):
        Z = int(problem["Z"])
        g = np.repeat(np.asarray(problem["g0"], dtype=np.complex128).reshape(1, Z), members, axis=0)
        f = np.repeat(np.asarray(problem["fields"][:n]).reshape(n, 1), members, axis=1)
        cport.line_steps(problem, g=g, fields=f, nsteps=n)
        return n * Z * members
    cport.sh_steps(problem, nsteps=n)
    return n * int(problem["L"]) * int(problem["R"])


def cpu_reference_rate(problem, seconds_target, min_steps=2, max_steps=400):
    """updates/s of the oracle C port on a bounded sample (first n time steps of the workload; LineMesh: one member per
    host thread, as the reference's process pool would run an ensemble)"""
    from oracle import cport

    members = cport.num_threads() if str(problem["kind"]).startswith("line") else 1
    t0 = time.perf_counter()
    cpu_steps(problem, 1, members)
    t1 = time.perf_counter() - t0
    n = int(max(min_steps, min(max_steps, seconds_target / max(t1, 1e-6))))
    t0 = time.perf_counter()
    upd = cpu_steps(problem, n, members)
    dt = time.perf_counter() - t0
    return upd / dt, n, dt, cport.num_threads()


def parity_prefix(sim, problem, batch, fields_all, n_check=32):
    """Self-check outside the timed region: n_check consecutive time steps of THIS workload around the pulse maximum, from a
    seeded state that populates every channel, on the schedule that is timed (CUDA graphs, fused kernels), compared with the
    oracle's C restatement (pinned to the reference by tests/golden).  -> the "parity" object of the bench line."""
    from oracle import cport

    is_line = str(problem["kind"]).startswith("line")
    n_all = len(problem["taus"])
    f2 = np.asarray(fields_all).reshape(n_all, -1)
    n = min(n_check, n_all)
    start = int(min(max(0, int(np.argmax(np.abs(f2[:, -1]))) - n // 2), n_all - n))
    taus = np.ascontiguousarray(problem["taus"][start : start + n])
    fields = np.ascontiguousarray(f2[start : start + n] if f2.shape[1] > 1 else f2[start : start + n, 0])
    rng = np.random.default_rng(11)
    if is_line:
        Z = int(problem["Z"])
        z = np.asarray(problem["z"])
        g0 = (rng.standard_normal(Z) + 1j * rng.standard_normal(Z)) * np.exp(-((z / z[-1]) ** 2) * 2)
        g0 = (g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(problem["delta_z"]))).reshape(1, Z)
        shape = (batch, 1, Z)
    else:
        L, R = int(problem["L"]), int(problem["R"])
        r = np.asarray(problem["r"])
        g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((r / r[-1]) ** 2) * 3)[None, :]
        g0 *= np.exp(-np.arange(L) / (L / 4))[:, None]
        g0 = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(problem["delta_r"]))
        shape = (batch, L, R)
    sim.write_g(np.ascontiguousarray(np.broadcast_to(g0, shape)))
    sim.step(taus, fields)
    g = sim.read_g()
    members = sorted({0, batch // 2, batch - 1})
    worst = 0.0
    for b in members:
        q = dict(problem)
        q["g0"], q["taus"] = (g0[0] if is_line else g0), taus
        q["fields"] = np.ascontiguousarray(fields[:, b]) if np.ndim(fields) == 2 else fields
        ref = cport.line_steps(q) if is_line else cport.sh_steps(q)
        got = g[b, 0] if is_line else g[b]
        worst = max(worst, float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
    return {"max_rel_err": worst, "steps": n, "first_step": start, "members_checked": members, "against": "oracle C port (oracle/c/restate.c), seeded state over all channels",
            "tolerance": 1e-10, "ok": bool(worst <= 1e-10)}


# ---------------------------------------------------------------------------------------------
# extra.* : the two real multi-GPU partitionings of BASELINE.json (configs[3] and configs[4]), measured in the same process
# group as the headline (VERDICT r01 "next round" 3).  Bounded: a slice of the pulse around its maximum; rates are per time step.
# ---------------------------------------------------------------------------------------------
def _max_over_ranks(ms, distributed):
    if not distributed:
        return float(ms)
    import torch
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def e2e_with_setup(local_rank):
    """configs[2] end to end through the PUBLIC API, set-up included: SphericalHarmonicSpecification(...)
    .to_sim() builds the time grid, field series, states, mask and Hamiltonian vectors on the host, run() evolves the whole pulse
    on the device and fills the datastores at the first and last time (store_data_every = -1, as scans do)."""
    import ionization_b200 as ion
    from ionization_b200 import potentials as P
    from ionization_b200 import states as S
    from ionization_b200 import units as u

    pw, rb = 200 * u.asec, 250 * u.bohr_radius
    t0 = time.perf_counter()
    spec = ion.mesh.SphericalHarmonicSpecification(
        "bench_c3_vel", r_bound=rb, r_points=2000, l_bound=500, time_initial=-5 * pw, time_final=5 * pw, time_step=1 * u.asec,
        electric_potential=P.SincPulse(pulse_width=pw, fluence=1 * u.Jcm2, phase=0, window=P.LogisticWindow(window_time=4 * pw, window_width=0.2 * pw)),
        use_numeric_eigenstates=False, test_states=[S.HydrogenBoundState(n, l) for n in range(1, 4) for l in range(n)],
        mask=P.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb, smoothness=8), operators=ion.mesh.SphericalHarmonicVelocityGaugeOperators(),
        evolution_method=ion.mesh.SplitInteractionOperator(), store_data_every=-1, device=local_rank,
    )
    sim = spec.to_sim()
    t1 = time.perf_counter()
    sim.run()
    norm = float(sim.data.norm[-1])
    t2 = time.perf_counter()
    updates = (sim.time_steps - 1) * 2000 * 500
    return {"value": updates / (t2 - t0), "unit": "updates/s", "setup_s": t1 - t0, "run_s": t2 - t1, "final_norm": norm,
            "api": "ionization_b200.mesh.SphericalHarmonicSpecification(...).to_sim().run(), wall clock, first call in the process (LU factors + graph capture included)"}


def extra_observed_every_step(problem, local_rank, stream, n_t=256):
    """store_data_every = 1 (the reference's default, mesh/sims.py:471): norm, inner products and norm by l after EVERY step.  On the
    fused schedule the reductions ride inside the step kernels (k_slab<OBS> / k_unit<LEN_STEP_OBS>); the ratio to the unobserved step is the
    cost of observing (north_star 4)."""
    import torch

    from ionization_b200 import _native as nat
    from ionization_b200 import engine

    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_NORM_BY_L
    n_t = min(n_t, len(problem["taus"]))
    start = max(0, len(problem["taus"]) // 2 - n_t // 2)
    taus, fields = problem["taus"][start : start + n_t], problem["fields"][start : start + n_t]
    out = {"time_steps": n_t, "observables": "norm, inner products with the test states, norm by l"}
    with engine.DeviceSimulation.from_problem(problem, device=local_rank) as sim:
        sim.set_stream(stream.cuda_stream)
        for name, mask in (("unobserved", np.zeros(n_t, dtype=np.uint8)), ("every_step", np.ones(n_t, dtype=np.uint8))):
            sim.run(taus, fields, mask, what)  # warm-up: graph capture for this pattern
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rec = sim.run(taus, fields, mask, what)
            e1.record(stream)
            e1.synchronize()
            out[f"us_per_time_step_{name}"] = 1e3 * e0.elapsed_time(e1) / n_t
            if name == "every_step":
                out["records"] = int(rec.shape[0])
                out["final_norm"] = float(rec[-1, 0, 0])
    out["ratio"] = out["us_per_time_step_every_step"] / out["us_per_time_step_unobserved"]
    return out


def _all_ok(ok, distributed):
    """collective vote: did every rank get through its (collective-free) set-up?  Keeps the ranks in step when one of them fails."""
    if not distributed:
        return bool(ok)
    import torch
    import torch.distributed as dist

    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t[0]))


def extra_c4_scan(rank, world, local_rank, stream, n_t=200, n_fluence=64, n_phase=64):
    """configs[3]: the 4096-member fluence x CEP scan (ionization_scans/scan_mesh.py:40-75) split as 4096 / N members per rank
    (strong scaling, no data-path collective), per-member norm and ionization fraction gathered inside the timed region
    (scan_utils.py:638-663 returns each finished sim to the submitter)."""
    import torch

    from ionization_b200 import _native as nat
    from ionization_b200 import coefficients as C
    from ionization_b200 import configs, engine, parallel
    from ionization_b200 import units as u

    distributed = world > 1
    p = configs.config4_member("LEN")
    L, R = int(p["L"]), int(p["R"])
    total = n_fluence * n_phase
    b0, b1 = parallel.shard_range(total, rank, world)
    nb = b1 - b0
    flu = np.geomspace(0.01, 20, n_fluence) * u.Jcm2
    ph = np.linspace(0, u.twopi, n_phase, endpoint=False)
    start = 1000 - n_t // 2
    times = p["times"][start : start + n_t + 1]
    pulses = [configs.sinc_pulse(200 * u.asec, flu[m // n_phase], ph[m % n_phase]) for m in range(b0, b1)]
    fields = C.field_series_batch("sh_len_so", pulses, times, p["time_step"], device=local_rank)  # E(t + dt/2) of the rank's members: one kernel (csrc/fields.cuh)
    taus = np.ascontiguousarray(p["taus"][start : start + n_t])
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS
    mask = np.zeros(n_t, dtype=np.uint8)
    mask[-1] = 1
    bound = np.asarray(p["state_bound"]).astype(bool)
    sim, err = None, None
    try:
        sim = engine.DeviceSimulation.from_problem(p, batch=nb, device=local_rank)
        sim.set_stream(stream.cuda_stream)
        sim.run(taus, fields, mask, what)  # warm-up: LU factors, graph capture
        sim.write_g_broadcast(p["g0"])
    except Exception as exc:  # noqa: BLE001
        err = f"{type(exc).__name__}: {exc}"
    if not _all_ok(err is None, distributed):
        if sim is not None:
            sim.close()
        return {"error": err or "set-up failed on another rank"}
    with sim:
        torch.cuda.synchronize()
        l0 = sim.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        rec = sim.run(taus, fields, mask, what)[0]  # [nb, 1 + 2 n_states]
        e1.record(stream)
        e1.synchronize()
        ips = rec[:, 1:].reshape(nb, -1, 2)
        ov = ips[..., 0] ** 2 + ips[..., 1] ** 2
        mine = {"first_member": b0, "norm": rec[:, 0].copy(), "ionization_fraction": 1.0 - ov[:, bound].sum(axis=1)}
        gathered = parallel.gather_objects(mine) if distributed else [mine]
        wall = time.perf_counter() - t0
        ms = _max_over_ranks(e0.elapsed_time(e1), distributed)
        launches = sim.launch_count - l0
    norms = np.concatenate([g["norm"] for g in sorted(gathered, key=lambda g: g["first_member"])])
    peak, _ = measured_peak_gbs()
    rate = total * L * R * n_t / (ms * 1e-3)
    return {
        "workload": f"configs[3]: {total}-member fluence x CEP scan, SphericalHarmonicMesh 1000x200 LEN split-operator; {n_t} of 2000 time steps around the pulse maximum",
        "members": total, "members_per_rank": nb, "time_steps": n_t, "ms_per_step": ms / n_t, "updates_per_s": rate, "scaling": "strong",
        "hbm_roofline_frac_per_gpu": rate / world * BYTES_PER_UPDATE / (peak * 1e9), "gather_wall_s_incl_run": wall, "members_gathered": int(len(norms)),
        "norm_min": float(norms.min()), "norm_max": float(norms.max()), "gpu_launches_per_rank": int(launches), "collective": "none on the data path; all_gather_object of per-member scalars at the end",
    }


def extra_c5_sharded(rank, world, local_rank, n_t=100):
    """configs[4]: one SphericalHarmonicMesh simulation r_points=16384, l_bound=4096 (length gauge), l-block sharded over the
    ranks with the engine's peer-memory halo exchange inside the captured step loop (csrc/halo.cuh).  At N = 1: the
    unsharded run.  Parity: the gathered shards against the unsharded run on rank 0, same steps."""
    import torch

    from ionization_b200 import _native as nat
    from ionization_b200 import configs, engine, parallel
    from ionization_b200 import units as u

    distributed = world > 1
    R, L = 16384, 4096
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=n_t,
                                           pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n_t / 2 * u.asec, time_final=n_t / 2 * u.asec)
    rng = np.random.default_rng(5)
    g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2) * 3)[None, :]
    g0 *= np.exp(-np.arange(L) / (L / 4))[:, None]
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))
    del g0
    out = {"workload": f"configs[4]: SphericalHarmonicMesh r_points={R} l_bound={L} LEN split-operator, {n_t} time steps, one simulation", "time_steps": n_t, "scaling": "strong"}
    peak, _ = measured_peak_gbs()

    def unsharded(env):
        keep = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            with engine.DeviceSimulation.from_problem(p, device=local_rank) as sim:
                st = torch.cuda.Stream()
                sim.set_stream(st.cuda_stream)
                sim.step(p["taus"], p["fields"])
                sim.write_g_broadcast(p["g0"])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                sim.step(p["taus"], p["fields"])
                e1.record(st)
                e1.synchronize()
                return e0.elapsed_time(e1) / n_t, sim.read_g()[0]
        finally:
            for k, v in keep.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)

    if not distributed:
        ms, _ = unsharded({})
        out.update({"ms_per_step": ms, "updates_per_s": R * L / (ms * 1e-3), "hbm_roofline_frac_per_gpu": R * L / (ms * 1e-3) * BYTES_PER_UPDATE / (peak * 1e9),
                    "partitioning": "unsharded (N = 1): r-segmented kernels, folded length-gauge step"})
        return out
    import torch.distributed as dist

    shard, err = None, None
    try:
        shard = parallel.ShardedSimulation(p, rank, world, device=local_rank)
    except Exception as exc:  # noqa: BLE001
        err = f"{type(exc).__name__}: {exc}"
    if not _all_ok(err is None, True):
        if shard is not None:
            shard.close()
        return {"error": err or "set-up failed on another rank"}
    shard.attach_peers()
    shard.step_device(p["taus"], p["fields"])  # warm-up: LU factors, graph capture
    shard.engine.synchronize()
    shard.engine.write_g(np.asarray(p["g0"])[shard.l_begin : shard.l_begin + shard.L].reshape(1, shard.L, R))
    dist.barrier()
    torch.cuda.synchronize()
    l0 = shard.engine.launch_count
    n_ex0, _ = shard.engine.halo_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(shard.stream)
    shard.step_device(p["taus"], p["fields"])
    e1.record(shard.stream)
    e1.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1), True) / n_t
    launches = shard.engine.launch_count - l0
    n_ex, aborted = shard.engine.halo_status()
    rec = parallel.all_reduce_observation(shard.partial_observation(nat.OBS_NORM), device=local_rank)
    blocks = parallel.gather_objects((shard.l_begin, shard.read_g()))
    shard.close()
    out.update({"ms_per_step": ms, "updates_per_s": R * L / (ms * 1e-3), "hbm_roofline_frac_per_gpu": R * L / (ms * 1e-3) / world * BYTES_PER_UPDATE / (peak * 1e9),
                "partitioning": f"{world} contiguous l-blocks of ~{L // world} channels cut at odd channels (every odd pair local), one ghost channel per neighbour = the read-only "
                                "even-pair partner of the block's first / last pair; every shard runs the one-kernel folded step of the unsharded engine",
                "halo_bytes_per_exchange_per_neighbour": R * 16, "exchanges_per_step": (n_ex - n_ex0) / n_t, "exchanges_done": n_ex, "halo_aborted": bool(aborted),
                "transport": "fused into the step kernel (PROG_LEN_STEP_HALO): the boundary CTAs store their channel into the neighbour's memory over NVLink (CUDA IPC peer memory) in the epilogue "
                             "and read the ghost partner the neighbour's previous launch delivered in the prologue; the stand-alone exchange kernel (counted in exchanges_per_step) only "
                             "before the first and after the last step of a call; NCCL only for rendezvous and the scalar all-reduce",
                "norm": float(rec[0]), "gpu_launches_per_rank": int(launches)})
    if rank == 0:
        ms_same, g_ref = unsharded({})  # the shards run the same folded one-kernel step as the unsharded engine
        ms_best = ms_same
        g = np.concatenate([blk for _, blk in sorted(blocks, key=lambda x: x[0])], axis=0)
        out.update({"max_rel_err_vs_unsharded": float(np.max(np.abs(g - g_ref)) / np.max(np.abs(g_ref))),
                    "unsharded_1gpu_ms_per_step_same_kernels": ms_same, "unsharded_1gpu_ms_per_step_best": ms_best,
                    "efficiency_vs_1gpu_same_kernels": ms_same / (world * ms), "efficiency_vs_1gpu_best": ms_best / (world * ms)})
    dist.barrier()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is one process that may use every host thread
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    wl = build_workload(args.workload)
    problem, desc = wl[0], wl[1]
    from oracle import cport

    is_line = str(problem["kind"]).startswith("line")
    L, R = (1, int(problem["Z"])) if is_line else (int(problem["L"]), int(problem["R"]))
    members = cport.num_threads() if is_line else 1
    # size the per-step sample so the whole run takes ~1-2 minutes
    t0 = time.perf_counter()
    cpu_steps(problem, 1, members)
    t1 = max(time.perf_counter() - t0, 1e-6)
    total_budget = 90.0
    n_t = int(max(1, min(200, total_budget / ((args.steps + args.warmup) * t1))))
    for _ in range(args.warmup):
        cpu_steps(problem, n_t, members)
    t0 = time.perf_counter()
    upd = 0
    for _ in range(args.steps):
        upd += cpu_steps(problem, n_t, members)
    dt = time.perf_counter() - t0
    value = upd / dt
    cores = cport.num_threads()
    sample = f"{n_t} of {len(problem['taus'])} time steps of the same mesh per bench step ({members} member(s))"
    line = {
        "impl": "reference", "metric": "grid-point updates/s (SphericalHarmonicMesh CN+split)", "value": value, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128 (f64)", "data": "synthetic", "config": {"workload": desc, "mesh_points": L * R, "time_steps_per_step": n_t},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample, "note": PORT_NOTE},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_vel")
    ap.add_argument("--time-steps", type=int, default=None, help="time steps per bench step (default: the workload's full pulse)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed self-check against the oracle")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.c4_scan / extra.c5_sharded (the two multi-GPU partitionings)")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    from ionization_b200 import engine
    from ionization_b200 import _native as nat

    if not torch.cuda.is_available() or engine.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: ionization_b200 has no CPU fallback")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        import torch.distributed as dist

        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    wl = build_workload(args.workload)
    problem, desc = wl[0], wl[1]
    batch = wl[2] if len(wl) > 2 else 1
    is_line = str(problem["kind"]).startswith("line")
    L, R = (1, int(problem["Z"])) if is_line else (int(problem["L"]), int(problem["R"]))
    n_t = len(problem["taus"]) if args.time_steps is None else min(args.time_steps, len(problem["taus"]))
    taus = np.ascontiguousarray(problem["taus"][:n_t])
    fields = np.ascontiguousarray((wl[3] if len(wl) > 3 else problem["fields"])[:n_t])
    updates_per_step = n_t * L * R * batch  # per GPU

    sim = engine.DeviceSimulation.from_problem(problem, batch=batch, device=local_rank)
    stream = torch.cuda.Stream()  # a capturable (non-legacy) stream: the engine replays CUDA graphs on it
    torch.cuda.set_stream(stream)
    sim.set_stream(stream.cuda_stream)
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS

    # pinned host buffers for the end-to-end path
    g0_full = np.array(np.broadcast_to(np.asarray(problem["g0"], dtype=np.complex128).reshape(1, L, R), (batch, L, R)), order="C", copy=True)
    g0_pinned = torch.from_numpy(g0_full.view(np.float64).reshape(batch, L, R, 2)).pin_memory()
    g0_np = g0_pinned.numpy().view(np.complex128).reshape(batch, L, R)
    gout_pinned = torch.empty((batch, L, R, 2), dtype=torch.float64).pin_memory()
    gout_np = gout_pinned.numpy().view(np.complex128).reshape(batch, L, R)
    del g0_full
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def reset():
        sim.write_g(g0_np)
        flush.zero_()  # L2 flush between timed iterations
        torch.cuda.synchronize()

    def timed(fn, k):
        """K steps, each bracketed by CUDA events on the launching stream; psi reset + L2 flush in between (untimed)"""
        total = 0.0
        for _ in range(k):
            reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total  # ms

    def step_resident():
        sim.step(taus, fields)

    def step_e2e():
        sim.write_g(g0_np)
        sim.step(taus, fields)
        sim.read_g(gout_np)
        sim.observe(what)

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        reset()
        step_resident()
    torch.cuda.synchronize()

    # ---- self-check against the oracle (untimed; rank 0) ----
    parity = None
    if rank == 0 and not args.no_parity and str(problem["kind"]) != "sh_len_adi":  # the C port has no ADI (tests compare ADI with oracle/restate.py)
        try:
            parity = parity_prefix(sim, problem, batch, wl[3] if len(wl) > 3 else problem["fields"])
        except Exception as exc:  # the checker is optional at bench time
            parity = {"max_rel_err": None, "ok": False, "error": str(exc)}

    # ---- timed: device-resident ----
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count
    wall0 = time.perf_counter()
    ms = timed(step_resident, args.steps)
    launches = sim.launch_count - launches0
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed: end-to-end through host buffers ----
    for _ in range(2):
        step_e2e()
    barrier()
    ms_e2e = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_e2e()
        e1.record(stream)
        e1.synchronize()
        ms_e2e += e0.elapsed_time(e1)
    barrier()

    if distributed:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel: CUDA events around every launch over a slice of the same steps ----
    roofline = None
    prof = {}
    if rank == 0:
        reset()
        n_prof = min(n_t, 200)
        prof = sim.profile(taus[:n_prof], fields[:n_prof])
        peak, peak_src = measured_peak_gbs()
        if prof:
            dom = max(prof, key=lambda k: prof[k][0])
            dom_ms, dom_n = prof[dom]
            total_ms = sum(v[0] for v in prof.values())
            steps_per_launch = 1.0  # every k_unit / k_slab launch streams the whole psi once
            alg_bytes = BYTES_PER_UPDATE * L * R * batch * steps_per_launch
            achieved = alg_bytes / (dom_ms / dom_n * 1e-3) / 1e9
            kname = {"slab": "k_slab", "adi_l": "k_adi_l", "len_ens": "k_len_ens"}.get(dom, f"k_unit<{dom}>")
            roofline = {
                "bound": "hbm", "kernel": kname, "time_steps_per_launch": steps_per_launch, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload, kname), "traffic_warm_l2": ncu_traffic_warm(args.workload, kname),
                "peak_source": peak_src, "avg_launch_us": 1e3 * dom_ms / dom_n, "share_of_step": dom_ms / total_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernels_us": {k: round(1e3 * v[0] / v[1], 3) for k, v in prof.items()},
                "kernel_launches_per_time_step": {k: v[1] / n_prof for k, v in prof.items()},
            }
            # second bound: the FP64 pipe.  Peak measured live (ion_fp64_peak: independent DFMA chains at
            # full occupancy); FP64 instructions per grid point and time step from the committed ncu capture of this workload
            # (profiles/ncu_traffic.json: sm__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on of one launch of each
            # kernel of a step), else the static count stated in DESIGN.md.
            try:
                fma_per_s = engine.fp64_peak(local_rank)
                ipp, ipp_src = fp64_inst_per_point(args.workload, prof, n_prof, L * R * batch)
                step_s = ms * 1e-3 / args.steps / n_t
                roofline["fp64"] = {
                    "peak_measured_tflops": 2e-12 * fma_per_s, "peak_fma_per_s": fma_per_s, "inst_per_point": ipp, "inst_per_point_source": ipp_src,
                    "achieved_inst_per_s": ipp * L * R * batch / step_s if ipp else None,
                    "frac": (ipp * L * R * batch / step_s / fma_per_s) if ipp else None,
                    "floor_us_per_time_step": (1e6 * ipp * L * R * batch / fma_per_s) if ipp else None,
                }
            except Exception as exc:
                roofline["fp64"] = {"error": str(exc)}

    value = world * args.steps * updates_per_step / (ms * 1e-3)
    value_e2e = world * args.steps * updates_per_step / (ms_e2e * 1e-3)
    peak, peak_src = measured_peak_gbs()
    n_states = len(problem["state_rows"])
    h2d = batch * L * R * 16 + n_t * batch * 8
    d2h = batch * L * R * 16 + batch * 8 * (1 + 2 * n_states)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and str(problem["kind"]) != "sh_len_adi":  # N = 1 only; the C port has no ADI
        try:
            v, n_cpu, dt_cpu, cores = cpu_reference_rate(problem, seconds_target=12.0)
            cpu_baseline = {"value": v, "unit": "updates/s", "cores": cores, "kind": "port",
                            "sample": f"first {n_cpu} of {len(problem['taus'])} time steps of the same mesh ({dt_cpu:.1f} s); oracle C restatement, OpenMP", "note": PORT_NOTE}
        except Exception as exc:  # the checker is optional at bench time
            cpu_baseline = {"value": None, "unit": "updates/s", "cores": 0, "kind": "port", "sample": f"unavailable: {exc}"}

    sim.close()
    with_setup = None
    if rank == 0 and args.workload == "c3_vel" and not args.no_extras:
        try:
            with_setup = e2e_with_setup(local_rank)
        except Exception as exc:  # noqa: BLE001
            with_setup = {"error": f"{type(exc).__name__}: {exc}"}
    extra = None
    if args.workload == "c3_vel" and not args.no_extras:
        extra = {}
        if rank == 0:
            try:
                extra["observed_every_step"] = extra_observed_every_step(problem, local_rank, stream)
            except Exception as exc:  # noqa: BLE001
                extra["observed_every_step"] = {"error": f"{type(exc).__name__}: {exc}"}
        for name, fn in (("c4_scan", lambda: extra_c4_scan(rank, world, local_rank, stream)), ("c5_sharded", lambda: extra_c5_sharded(rank, world, local_rank))):
            try:
                extra[name] = fn()
            except Exception as exc:  # an extra must never take the headline down with it
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
                if distributed:  # the ranks may be out of step now: no further collectives
                    break
    if rank == 0:
        line = {
            "metric": "grid-point updates/s (SphericalHarmonicMesh CN+split)", "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "complex128 (f64)", "data": "synthetic",
            "config": {"workload": desc, "mesh_points": L * R, "time_steps_per_step": n_t, "sims_per_gpu": batch, "parallelism": f"ensemble x{world} (independent sims, no collective)",
                       "l2": "flushed between timed iterations (256 MB write); psi (16 B/pt) + CN factors (16 B/pt) are L2-resident within an iteration by design"},
            "hbm_roofline_frac_step": value / world * BYTES_PER_UPDATE / (peak * 1e9), "us_per_time_step": 1e3 * ms / args.steps / n_t,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "e2e": {"value": value_e2e, "unit": "updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps, "with_setup": with_setup},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
