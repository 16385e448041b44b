"""GPU tests of (1) l-block sharding -- several shards on ONE GPU with an in-process halo exchange must reproduce the
unsharded run and the reference fixture -- and (2) r-segmented kernels (r_points > 4096) against the oracle."""
import numpy as np
import pytest

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _run_sharded(p, world, n_steps=None, what=0, cut_parity=None):
    from ionization_b200 import parallel

    shards = [parallel.ShardedSimulation(p, r, world, device=0, use_torch_stream=False, cut_parity=cut_parity) for r in range(world)]
    ex = parallel.LocalExchanger(shards)
    taus, fields = p["taus"][:n_steps], p["fields"][:n_steps]
    import torch

    for tau, f in zip(taus, fields):
        for ph in range(shards[0].n_phases):
            if ph in shards[0].halo_phases:
                for s in shards:
                    s.engine.synchronize()
                ex.exchange_all()
                torch.cuda.synchronize()
            for s in shards:
                s.run_phase(ph, tau, f)
    for s in shards:
        s.engine.synchronize()
    g = np.concatenate([s.read_g() for s in shards], axis=0)
    rec = None
    if what:
        ex.exchange_all()
        torch.cuda.synchronize()
        recs = [s.partial_observation(what) for s in shards]
        rec = parallel.combine_observations(recs, what, n_states=len(p["state_l"]), l_counts=[s.L for s in shards])
    for s in shards:
        s.close()
    return g, rec


@pytest.mark.parametrize("name", ["sh_len_so_100x10", "sh_vel_so_60x8", "sh_len_so_datastores_120x12", "sh_vel_so_datastores_120x12"])
@pytest.mark.parametrize("world", [2, 3])
def test_l_block_shards_reproduce_reference(name, world):
    from ionization_b200 import _native as nat

    p = load_golden(name)
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_NORM_BY_L | nat.OBS_R | nat.OBS_Z | nat.OBS_H0
    g, rec = _run_sharded(p, world, what=what)
    assert rel_err(g, p["g_final"]) < TOL
    ns = len(p["state_l"])
    assert abs(rec[0] - p["norm"][-1]) < TOL
    ips = rec[1 : 1 + 2 * ns].reshape(ns, 2)
    assert np.max(np.abs(ips[:, 0] + 1j * ips[:, 1] - p["inner_products"][-1])) < TOL
    if "norm_by_l" in p:
        L = int(p["L"])
        c = 1 + 2 * ns
        assert np.max(np.abs(rec[c : c + L] - p["norm_by_l"][-1])) < TOL
        assert abs(rec[c + L] - p["r_expectation"][-1]) < TOL * abs(p["r_expectation"][-1])
        assert abs(rec[c + L + 1] - p["z_expectation"][-1]) < TOL * abs(p["r_expectation"][-1])
        assert abs(rec[c + L + 2] - p["internal_energy"][-1]) < TOL * abs(p["internal_energy"][-1])


@pytest.mark.parametrize("kind", ["LEN", "VEL"])
def test_l_block_shards_equal_unsharded_run_on_a_larger_mesh(kind):
    from ionization_b200 import configs, engine
    from ionization_b200 import units as u

    p = configs.spherical_harmonic_problem(r_bound=60 * u.bohr_radius, r_points=600, l_bound=64, gauge=kind, n_steps=40,
                                           pulse=configs.sinc_pulse(20 * u.asec, 5 * u.Jcm2), time_initial=-20 * u.asec, time_final=20 * u.asec)
    with engine.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g_ref = sim.read_g()[0]
    g, _ = _run_sharded(p, 4)
    assert rel_err(g, g_ref) < 1e-12


@pytest.mark.parametrize("cut_parity", [0, 1])
@pytest.mark.parametrize("transport", ["phases", "device"])
@pytest.mark.parametrize("world", [2, 3, 4])
def test_length_gauge_shards_with_even_and_odd_cuts(cut_parity, transport, world):
    """length gauge: blocks cut at even channels (the straddling odd pairs are evaluated on both sides, single-sweep kernels) and
    at odd channels (every odd pair local; linked shards run the one-kernel folded step with the ghost channels as the read-only
    even-pair partners) against the reference fixture, observables included"""
    from ionization_b200 import _native as nat

    p = load_golden("sh_len_so_datastores_120x12")
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_NORM_BY_L | nat.OBS_R | nat.OBS_Z
    run = _run_sharded if transport == "phases" else _run_sharded_device
    g, rec = run(p, world, what=what, cut_parity=cut_parity)
    assert rel_err(g, p["g_final"]) < TOL
    ns, L = len(p["state_l"]), int(p["L"])
    assert abs(rec[0] - p["norm"][-1]) < TOL
    c = 1 + 2 * ns
    assert np.max(np.abs(rec[c : c + L] - p["norm_by_l"][-1])) < TOL
    assert abs(rec[c + L] - p["r_expectation"][-1]) < TOL * abs(p["r_expectation"][-1])
    assert abs(rec[c + L + 1] - p["z_expectation"][-1]) < TOL * abs(p["r_expectation"][-1])


def test_odd_cuts_are_refused_for_the_velocity_gauge():
    from ionization_b200 import engine, exceptions

    with pytest.raises(exceptions.IonizationException):
        engine.DeviceSimulation("sh_vel_so", 4, 64, batch=1, device=0, L_total=12, l_begin=3)


@pytest.mark.parametrize("kind", ["LEN", "VEL"])
@pytest.mark.parametrize("R", [5000, 4609])
def test_r_segmented_kernels_match_oracle(kind, R):
    """r_points > 4096: every channel is processed by several CTAs with recomputed halos (kernels.cuh)"""
    from ionization_b200 import configs, engine
    from ionization_b200 import units as u
    from oracle import cport

    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=6, gauge=kind, n_steps=30,
                                           pulse=configs.sinc_pulse(20 * u.asec, 5 * u.Jcm2), time_initial=-15 * u.asec, time_final=15 * u.asec)
    # spread the wavefunction over the whole radial range so that every segment boundary carries amplitude
    rng = np.random.default_rng(R)
    g0 = (rng.standard_normal((6, R)) + 1j * rng.standard_normal((6, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2))[None, :]
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))
    ref = cport.sh_steps(p)
    with engine.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0]
        norm = sim.observe(engine.nat.OBS_NORM)[0, 0]
    assert rel_err(g, ref) < TOL
    assert abs(norm - np.sum(np.abs(ref) ** 2) * float(p["delta_r"])) < TOL


def test_segmented_line_split_operator_long_mesh():
    """LineMesh with 2^14 / 2^16 points (SO both gauges, CN) against the C oracle"""
    from ionization_b200 import engine
    from oracle import cport

    for kind, Z in (("line_len_so", 2 ** 14), ("line_vel_so", 2 ** 14), ("line_len_cn", 2 ** 14), ("line_len_cn", 2 ** 16)):
        p = dict(load_golden(f"{kind}_1024"))
        z = np.linspace(-1, 1, Z) * p["z"][-1] * (Z // 1024)
        dz = z[1] - z[0]
        scale = (float(p["delta_z"]) / dz) ** 2
        p.update(Z=Z, z=z, delta_z=dz, h_off=np.full(Z - 1, p["h_off"][0] * scale), w_z=z * (p["w_z"][-1] / p["z"][-1]), mask=np.ones(Z),
                 v_pref=float(p["v_pref"]) * float(p["delta_z"]) / dz)
        p["h_diag"] = np.full(Z, -2 * p["h_off"][0]) + 0j + np.interp(z, load_golden(f"{kind}_1024")["z"], np.real(load_golden(f"{kind}_1024")["h_diag"]) + 2 * load_golden(f"{kind}_1024")["h_off"][0])
        rng = np.random.default_rng(1)
        g0 = (rng.standard_normal(Z) + 1j * rng.standard_normal(Z)) * np.exp(-((z / z[-1]) ** 2) * 2)
        p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * dz)
        p["state_rows"] = p["g0"][None, :]
        p["fields"] = p["fields"] * 0.05
        ref = cport.line_steps(p)
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"], p["fields"])
            g = sim.read_g()[0, 0]
        assert rel_err(g, ref) < TOL, kind


def _run_sharded_device(p, world, n_steps=None, what=0, cut_parity=None, devices=None):
    """the production transport: shards linked with ion_sim_attach_peer, advanced by the device-resident loop with the
    engine's own halo-exchange kernel (flags + stores into the neighbour's ghost channel).  Several shards in ONE process
    on one GPU: every shard is driven from its own host thread, as every rank would be from its own process."""
    import threading

    from ionization_b200 import parallel

    shards = [parallel.ShardedSimulation(p, r, world, device=devices[r] if devices else 0, use_torch_stream=False, cut_parity=cut_parity) for r in range(world)]
    parallel.ShardedSimulation.attach_local(shards)
    taus, fields = p["taus"][:n_steps], p["fields"][:n_steps]
    for s in shards:
        s.engine.prepare(float(taus[0]))  # everything that allocates / frees happens before the first hand-shake
    errors = []

    def drive(s):
        try:
            s.step_device(taus, fields)
            if what:
                s.engine.exchange_halos()
            s.engine.synchronize()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=drive, args=(s,)) for s in shards]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for s in shards:
        n_ex, aborted = s.engine.halo_status()
        assert not aborted and n_ex > 0
    g = np.concatenate([s.read_g() for s in shards], axis=0)
    rec = None
    if what:
        recs = [s.engine.observe(what)[0] for s in shards]
        rec = parallel.combine_observations(recs, what, n_states=len(p["state_l"]), l_counts=[s.L for s in shards])
    for s in shards:
        s.close()
    return g, rec


@pytest.mark.parametrize("name", ["sh_len_so_100x10", "sh_vel_so_60x8", "sh_vel_so_datastores_120x12"])
@pytest.mark.parametrize("world", [2, 3])
def test_peer_memory_halo_exchange_reproduces_reference(name, world):
    from ionization_b200 import _native as nat

    p = load_golden(name)
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_Z
    g, rec = _run_sharded_device(p, world, what=what)
    assert rel_err(g, p["g_final"]) < TOL
    ns = len(p["state_l"])
    assert abs(rec[0] - p["norm"][-1]) < TOL
    if "z_expectation" in p:
        assert abs(rec[1 + 2 * ns] - p["z_expectation"][-1]) < TOL * abs(p["r_expectation"][-1])


@pytest.mark.parametrize("kind", ["LEN", "VEL"])
def test_peer_memory_halo_exchange_equals_unsharded_run_on_a_larger_mesh(kind):
    from ionization_b200 import configs, engine
    from ionization_b200 import units as u

    p = configs.spherical_harmonic_problem(r_bound=60 * u.bohr_radius, r_points=600, l_bound=64, gauge=kind, n_steps=70,
                                           pulse=configs.sinc_pulse(20 * u.asec, 5 * u.Jcm2), time_initial=-35 * u.asec, time_final=35 * u.asec)
    with engine.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g_ref = sim.read_g()[0]
    g, _ = _run_sharded_device(p, 4)
    assert rel_err(g, g_ref) < 1e-12


@pytest.mark.parametrize("shape", [(600, 64, 70), (600, 64, 150), (2000, 24, 70), (3000, 16, 70), (5000, 16, 70)])
def test_halo_exchange_fused_into_the_length_gauge_step_equals_unsharded_run(shape):
    """PROG_LEN_STEP_HALO: the boundary CTAs of the folded step store their channel into the neighbour's memory and read their ghost
    partner from the slot the neighbour's previous launch filled -- no exchange kernel between fused steps.  The default for shards
    whose neighbours live on OTHER GPUs (a CTA that spins for a neighbour's kernel must not share a device with it), so this test
    needs two GPUs: one shard per device, graph chunks of 64 steps, r-segments included."""
    import torch

    from ionization_b200 import configs, engine
    from ionization_b200 import units as u

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs two GPUs (one l-block shard per device)")
    world = min(n_dev, 4)
    R, L, n = shape
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=n,
                                           pulse=configs.sinc_pulse(20 * u.asec, 5 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
    for steps in (n, 1):  # a single step: the stand-alone exchange only
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"][:steps], p["fields"][:steps])
            g_ref = sim.read_g()[0]
        g, _ = _run_sharded_device(p, world, n_steps=steps, devices=list(range(world)))
        assert rel_err(g, g_ref) < 1e-12


@pytest.mark.parametrize("kind", ["LEN", "VEL"])
def test_r_segmented_kernels_with_more_ctas_than_fit_on_the_gpu(kind, monkeypatch):
    """Regression: the halo rows of one segment's CTA are the interior rows of its neighbour's, so a segmented kernel that
    ran in place gave wrong results as soon as the grid needed more than one wave of CTAs (a later CTA read rows its
    neighbour had already advanced).  8192 rows x 64 channels = 6 segments x 33 units = 198 CTAs of 448/512 threads on 148
    SMs; both the fused single-GPU schedule and the single-sweep schedule the shards use are checked against the oracle."""
    from ionization_b200 import configs, engine
    from ionization_b200 import units as u
    from oracle import cport

    R, L, n = 8192, 64, 24
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge=kind, n_steps=n,
                                           pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
    rng = np.random.default_rng(7)
    g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2) * 3)[None, :]
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))
    ref = cport.sh_steps(p)
    for env in ({}, {"ION_NO_LEN_FOLD": "1", "ION_NO_SLAB": "1"}, {"ION_NO_LEN_FOLD": "1", "ION_NO_SLAB": "1", "ION_NO_PDL": "1", "ION_NO_GRAPHS": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"], p["fields"])
            g = sim.read_g()[0]
        for k in env:
            monkeypatch.delenv(k)
        assert rel_err(g, ref) < TOL, env


def test_segmented_line_crank_nicolson_ensemble_spans_several_waves():
    """LineMesh CN, 2^14 points (11 segments) x 48 members = 528 CTAs: every member must equal its single-member run"""
    from ionization_b200 import engine

    base = dict(load_golden("line_len_cn_1024"))
    Z = 2 ** 14
    z = np.linspace(-1, 1, Z) * base["z"][-1] * (Z // 1024)
    dz = z[1] - z[0]
    scale = (float(base["delta_z"]) / dz) ** 2
    p = dict(base)
    p.update(Z=Z, z=z, delta_z=dz, h_off=np.full(Z - 1, base["h_off"][0] * scale), w_z=z * (base["w_z"][-1] / base["z"][-1]), mask=np.ones(Z))
    p["h_diag"] = np.full(Z, -2 * p["h_off"][0]) + 0j + np.interp(z, base["z"], np.real(base["h_diag"]) + 2 * base["h_off"][0])
    rng = np.random.default_rng(3)
    g0 = (rng.standard_normal(Z) + 1j * rng.standard_normal(Z)) * np.exp(-((z / z[-1]) ** 2) * 2)
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * dz)
    p["state_rows"] = p["g0"][None, :]
    n = 12
    members = 48
    fields = 0.05 * np.outer(base["fields"][:n], np.linspace(0.2, 1.5, members))
    with engine.DeviceSimulation.from_problem(p, batch=members) as sim:
        sim.step(p["taus"][:n], fields)
        g = sim.read_g()[:, 0]
    for b in (0, 17, members - 1):
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"][:n], np.ascontiguousarray(fields[:, b]))
            g1 = sim.read_g()[0, 0]
        assert rel_err(g[b], g1) < 1e-13


@pytest.mark.parametrize("Z", [4097, 2 ** 13 + 37])
def test_line_crank_nicolson_eight_rows_per_thread_equals_four_rows_per_thread(Z, monkeypatch):
    """Long LineMesh channels of the Crank-Nicolson program take eight rows per thread (layout with M = 8, 256-thread segment CTAs,
    32-thread halos); ION_LINE_M=4 keeps the four-row kernels.  Same inputs, both against each other and against the oracle."""
    from ionization_b200 import engine
    from oracle import cport

    base = dict(load_golden("line_len_cn_1024"))
    z = np.linspace(-1, 1, Z) * base["z"][-1] * (Z / 1024)  # ragged sizes: the last segment is partly padding
    dz = z[1] - z[0]
    scale = (float(base["delta_z"]) / dz) ** 2
    p = dict(base)
    p.update(Z=Z, z=z, delta_z=dz, h_off=np.full(Z - 1, base["h_off"][0] * scale), w_z=z * (base["w_z"][-1] / base["z"][-1]), mask=np.cos(np.linspace(0, 1.0, Z)) ** 0.125)
    p["h_diag"] = np.full(Z, -2 * p["h_off"][0]) + 0j + np.interp(z, base["z"], np.real(base["h_diag"]) + 2 * base["h_off"][0])
    rng = np.random.default_rng(11)
    g0 = (rng.standard_normal(Z) + 1j * rng.standard_normal(Z)) * np.exp(-((z / z[-1]) ** 2) * 2)
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * dz)
    p["state_rows"] = p["g0"][None, :]
    n = 10
    p["taus"], p["fields"] = p["taus"][:n], 0.05 * np.asarray(p["fields"][:n])
    ref = cport.line_steps(p)
    out = {}
    for m in ("8", "4"):
        monkeypatch.setenv("ION_LINE_M", m)
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"], p["fields"])
            out[m] = sim.read_g()[0, 0]
            norm = sim.observe(engine.nat.OBS_NORM)[0, 0]
        assert rel_err(out[m], ref) < TOL, m
        assert abs(norm - np.sum(np.abs(ref) ** 2) * dz) < TOL
    monkeypatch.delenv("ION_LINE_M")
    assert rel_err(out["8"], out["4"]) < 1e-12


def test_half_warp_halos_equal_full_warp_halos(monkeypatch):
    """Length-gauge r-segments recompute 16 halo threads (64 rows) per side when the host finds the LU multipliers below 1e-18 over any
    aligned 64 rows; ION_HALO16=0 keeps the 32-thread halos.  Same mesh, both ways, against each other and the oracle."""
    from ionization_b200 import configs, engine
    from ionization_b200 import units as u
    from oracle import cport

    R, L, n = 6000, 10, 16
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=n,
                                           pulse=configs.sinc_pulse(20 * u.asec, 20 * u.Jcm2), time_initial=-n / 2 * u.asec, time_final=n / 2 * u.asec)
    rng = np.random.default_rng(4)
    g0 = (rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))) * np.exp(-((p["r"] / p["r"][-1]) ** 2) * 3)[None, :]
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * float(p["delta_r"]))
    ref = cport.sh_steps(p)
    out = {}
    for h in ("1", "0"):
        monkeypatch.setenv("ION_HALO16", h)
        with engine.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"], p["fields"])
            out[h] = sim.read_g()[0]
        assert rel_err(out[h], ref) < TOL, h
    monkeypatch.delenv("ION_HALO16")
    assert rel_err(out["1"], out["0"]) < 1e-13
