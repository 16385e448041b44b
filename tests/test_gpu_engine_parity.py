"""GPU parity: the CUDA engine, called through the C-ABI, against fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, generator oracle/make_golden.py) and against the oracle on seeded inputs.

Tolerance (BASELINE.json north_star): <= 1e-10 relative on the complex128 wavefunction (max-norm), norm and
overlaps after the full pulse.  The engine typically lands at 1e-14..1e-13.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10

SH_SMALL = [
    "sh_len_so_100x10",
    "sh_len_so_101x11",
    "sh_len_so_100x11",
    "sh_len_so_101x10",
    "sh_vel_so_60x8",
    "sh_vel_so_61x9",
    "sh_vel_so_60x9",
    "sh_vel_so_61x8",
    "sh_len_so_datastores_120x12",
    "sh_vel_so_datastores_120x12",
    "sh_len_adi_64x8",
    "sh_len_adi_65x9",
]


def _engine():
    from ionization_b200 import engine

    return engine


@pytest.mark.parametrize("name", SH_SMALL)
def test_sh_final_wavefunction_matches_reference(name):
    eng = _engine()
    p = load_golden(name)
    with eng.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0]
    assert rel_err(g, p["g_final"]) < TOL


@pytest.mark.parametrize("name", SH_SMALL)
def test_sh_norm_and_inner_products_every_step(name):
    """store_data_every=1: observation after every step (unfused path) -- mesh/sims.py:290-294."""
    eng = _engine()
    p = load_golden(name)
    what = eng.nat.OBS_NORM | eng.nat.OBS_INNER_PRODUCTS
    n = len(p["taus"])
    with eng.DeviceSimulation.from_problem(p) as sim:
        rec0 = sim.observe(what)[0]
        rec = sim.run(p["taus"], p["fields"], np.ones(n, dtype=np.uint8), what)[:, 0, :]
        g = sim.read_g()[0]
    rec = np.concatenate([rec0[None, :], rec], axis=0)
    ns = len(p["state_l"])
    norm = rec[:, 0]
    ips = rec[:, 1 : 1 + 2 * ns].reshape(-1, ns, 2)
    ips = ips[..., 0] + 1j * ips[..., 1]
    assert np.max(np.abs(norm - p["norm"])) < TOL
    assert np.max(np.abs(ips - p["inner_products"])) < TOL
    assert rel_err(g, p["g_final"]) < TOL


@pytest.mark.parametrize("name", ["sh_len_so_100x10", "sh_vel_so_60x8"])
def test_fused_and_unfused_paths_agree(name):
    eng = _engine()
    p = load_golden(name)
    n = len(p["taus"])
    with eng.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g_fused = sim.read_g()[0]
    with eng.DeviceSimulation.from_problem(p) as sim:
        for k in range(n):
            sim.step(p["taus"][k : k + 1], p["fields"][k : k + 1])
        g_single = sim.read_g()[0]
    assert rel_err(g_fused, g_single) < 1e-12


@pytest.mark.parametrize("name", ["sh_len_so_datastores_120x12", "sh_vel_so_datastores_120x12"])
def test_all_observables(name):
    eng = _engine()
    nat = eng.nat
    p = load_golden(name)
    what = nat.OBS_NORM | nat.OBS_INNER_PRODUCTS | nat.OBS_NORM_BY_L | nat.OBS_R | nat.OBS_Z | nat.OBS_H0 | nat.OBS_NORM_WITHIN
    n = len(p["taus"])
    L = int(p["L"])
    ns = len(p["state_l"])
    radii = p["norm_within_radii"]
    with eng.DeviceSimulation.from_problem(p, radii=radii) as sim:
        rec0 = sim.observe(what)[0]
        rec = sim.run(p["taus"], p["fields"], np.ones(n, dtype=np.uint8), what)[:, 0, :]
    rec = np.concatenate([rec0[None, :], rec], axis=0)
    c = 0
    norm = rec[:, c]
    c += 1
    c += 2 * ns
    nbl = rec[:, c : c + L]
    c += L
    r_exp = rec[:, c]
    z_exp = rec[:, c + 1]
    h0 = rec[:, c + 2]
    within = rec[:, c + 3 : c + 3 + len(radii)]
    assert np.max(np.abs(norm - p["norm"])) < TOL
    assert np.max(np.abs(nbl - p["norm_by_l"])) < TOL
    assert rel_err(r_exp, p["r_expectation"]) < TOL
    assert np.max(np.abs(z_exp - p["z_expectation"])) < TOL * np.max(np.abs(p["r_expectation"]))
    assert rel_err(h0, p["internal_energy"]) < TOL
    assert np.max(np.abs(within - p["norm_within_radius"])) < TOL
    if "total_energy" in p:
        # <H> = <H0> + E(t + dt/2) * (-q) <z>   (mesh_operators.py:1008-1035)
        total = h0 + p["efield_half_at_data_times"] * (-float(p["test_charge"])) * z_exp
        assert rel_err(total, p["total_energy"]) < TOL


@pytest.mark.parametrize("name", ["c1_sh_len_so_500x50", "c1_sh_vel_so_500x50"])
def test_config1_full_pulse(name):
    """BASELINE.json configs[0]: hydrogen 1s, r_points=500, l_bound=50, Sinc pulse, 2000 steps."""
    eng = _engine()
    p = load_golden(name)
    what = eng.nat.OBS_NORM | eng.nat.OBS_INNER_PRODUCTS
    n = len(p["taus"])
    mask = np.zeros(n, dtype=np.uint8)
    mask[99::100] = 1  # store_data_every=100 -> data at time indices 100, 200, ... (0 handled separately)
    mask[-1] = 1
    with eng.DeviceSimulation.from_problem(p) as sim:
        rec0 = sim.observe(what)[0]
        rec = sim.run(p["taus"], p["fields"], mask, what)[:, 0, :]
        g = sim.read_g()[0]
    rec = np.concatenate([rec0[None, :], rec], axis=0)
    ns = len(p["state_l"])
    ips = rec[:, 1 : 1 + 2 * ns].reshape(-1, ns, 2)
    ips = ips[..., 0] + 1j * ips[..., 1]
    assert rel_err(g, p["g_final"]) < TOL
    assert np.max(np.abs(rec[:, 0] - p["norm"])) < TOL
    assert np.max(np.abs(ips - p["inner_products"])) < TOL
    # ionization fraction = 1 - bound_state_overlap[-1]  
    bound = p["state_bound"].astype(bool)
    ion_ref = 1 - np.sum(np.abs(p["inner_products"][-1][bound]) ** 2)
    ion_gpu = 1 - np.sum(np.abs(ips[-1][bound]) ** 2)
    assert abs(ion_gpu - ion_ref) <= TOL * abs(ion_ref)


@pytest.mark.parametrize("name", ["known_sh_len_so_500x200", "known_sh_vel_so_500x200", "known_sh_len_adi_500x200"])
def test_known_answers_of_the_reference(name):
    """dev/meshes/mesh_refactoring_helper.py:204-251: final initial-state overlaps 0.312928752359 (LEN SO),
    0.319513371899 (VEL SO) and 0.312910470190 (LEN ADI), numeric eigenstates, 800 steps."""
    eng = _engine()
    p = load_golden(name)
    expected = {"known_sh_len_so_500x200": 0.312928752359, "known_sh_vel_so_500x200": 0.319513371899,
                "known_sh_len_adi_500x200": 0.312910470190}[name]
    what = eng.nat.OBS_NORM | eng.nat.OBS_INNER_PRODUCTS
    with eng.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        rec = sim.observe(what)[0]
        g = sim.read_g()[0]
    ns = len(p["state_l"])
    ips = rec[1 : 1 + 2 * ns].reshape(ns, 2)
    ips = ips[:, 0] + 1j * ips[:, 1]
    overlap = abs(ips[int(p["initial_state_index"])]) ** 2
    assert abs(overlap - expected) < 5e-12  # 12 printed digits
    assert abs(overlap - float(p["initial_state_overlap_final"])) < TOL
    assert rel_err(g, p["g_final"]) < TOL


@pytest.mark.parametrize("name", ["line_len_so_1024", "line_len_so_1023", "line_vel_so_1024", "line_vel_so_1023", "line_len_cn_1024", "line_len_cn_1023"])
def test_line_mesh_programs(name):
    eng = _engine()
    p = load_golden(name)
    with eng.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0, 0]
    assert rel_err(g, p["g_final"]) < TOL


@pytest.mark.parametrize("name", ["known_line_len_cn_4096", "known_line_len_so_4096", "known_line_vel_so_4096"])
def test_line_known_answers_of_the_reference(name):
    """dev/meshes/mesh_refactoring_helper.py:40-63,:204-251: QHO driven by a sine wave, 4096 points, 1000 steps; final
    initial-state overlaps 0.370010185740 (LEN CN), 0.370008474418 (LEN SO), 0.370924310122 (VEL SO)"""
    eng = _engine()
    p = load_golden(name)
    expected = {"known_line_len_cn_4096": 0.370010185740, "known_line_len_so_4096": 0.370008474418, "known_line_vel_so_4096": 0.370924310122}[name]
    what = eng.nat.OBS_NORM | eng.nat.OBS_INNER_PRODUCTS
    with eng.DeviceSimulation.from_problem(p) as sim:
        sim.step(p["taus"], p["fields"])
        rec = sim.observe(what)[0]
        g = sim.read_g()[0, 0]
    ns = len(p["state_rows"])
    ips = rec[1 : 1 + 2 * ns].reshape(ns, 2)
    overlap = abs(ips[int(p["initial_state_index"]), 0] + 1j * ips[int(p["initial_state_index"]), 1]) ** 2
    assert abs(overlap - expected) < 5e-12
    assert rel_err(g, p["g_final"]) < TOL


def test_line_cn_ensemble_with_different_fields():
    """config 2 shape: a batch of LineMesh CN simulations with different pulses; every member has its own matrix"""
    eng = _engine()
    from oracle import cport

    p = load_golden("line_len_cn_1024")
    scales = np.array([1.0, -0.5, 3.0, 0.0, 10.0])
    fields = p["fields"][:, None] * scales[None, :]
    with eng.DeviceSimulation.from_problem(p, batch=len(scales)) as sim:
        sim.step(p["taus"], fields)
        gb = sim.read_g()[:, 0, :]
    assert rel_err(gb[0], p["g_final"]) < TOL
    g0 = np.repeat(np.asarray(p["g0"])[None, :], len(scales), axis=0)
    ref = cport.line_steps(p, g=g0, fields=fields)
    assert rel_err(gb, ref) < TOL


def test_ensemble_batch_matches_independent_runs():
    """scan ensemble (ionization_scans/scan_mesh.py:40-68): batch members with different field amplitudes evolve
    independently and identically to single runs."""
    eng = _engine()
    p = load_golden("sh_len_so_100x10")
    scales = np.array([1.0, 0.5, -0.25, 2.0])
    n = len(p["taus"])
    fields = p["fields"][:, None] * scales[None, :]
    with eng.DeviceSimulation.from_problem(p, batch=len(scales)) as sim:
        sim.step(p["taus"], fields)
        gb = sim.read_g()
    assert rel_err(gb[0], p["g_final"]) < TOL
    from oracle import restate

    for k, sc in enumerate(scales):
        q = dict(p)
        q["fields"] = p["fields"] * sc
        ref = restate.run_sh(q, store_every_step=False)["g"]
        assert rel_err(gb[k], ref) < TOL


def test_field_free_evolution_preserves_norm_and_overlaps():
    """tests/mesh/test_sims.py:31-82 of the reference: 100 field-free steps keep norm and overlaps (atol 1e-14 there,
    with numeric eigenstates; here analytic hydrogen states on the known-answer mesh with numeric eigenstates)."""
    eng = _engine()
    p = load_golden("known_sh_len_so_500x200")
    what = eng.nat.OBS_NORM | eng.nat.OBS_INNER_PRODUCTS
    n = 100
    with eng.DeviceSimulation.from_problem(p) as sim:
        rec0 = sim.observe(what)[0]
        sim.step(p["taus"][:n], np.zeros(n))
        rec1 = sim.observe(what)[0]
    ns = len(p["state_l"])
    ov0 = rec0[1 : 1 + 2 * ns].reshape(ns, 2)
    ov1 = rec1[1 : 1 + 2 * ns].reshape(ns, 2)
    assert abs(rec0[0] - rec1[0]) < 1e-13
    assert np.max(np.abs(np.sum(ov0**2, axis=1) - np.sum(ov1**2, axis=1))) < 1e-13


# ---------------------------------------------------------------------------------------------
# cy.tdma drop-in (tests/test_tdma.py:12-26 of the reference, seeded)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 3, 17, 500, 1000, 5000, 20000])
def test_tdma_agrees_with_oracle_and_solves_system(n):
    from scipy import sparse

    from oracle import restate

    eng = _engine()
    rng = np.random.default_rng(n)
    crs = lambda k: rng.random(k) + 1j * rng.random(k)  # tests/conftest.py:4-5 of the reference
    a, b, c, d = crs(n - 1), crs(n) + 2.0, crs(n - 1), crs(n)
    dia = sparse.diags([a, b, c], offsets=[-1, 0, 1]).todia()
    x = eng.tdma(dia, d)
    x_ref = restate.tdma(a, b, c, d)
    assert np.allclose(x, x_ref, rtol=1e-11, atol=1e-13)
    assert np.allclose(dia.dot(x), d)
    if n <= 1000:
        assert np.allclose(x, np.linalg.inv(dia.toarray()).dot(d))


def test_tdma_batched():
    from oracle import restate

    eng = _engine()
    rng = np.random.default_rng(0)
    B, n = 37, 129
    crs = lambda *s: rng.random(s) + 1j * rng.random(s)
    a, b, c, d = crs(B, n - 1), crs(B, n) + 2.0, crs(B, n - 1), crs(B, n)
    x = eng.tdma((a, b, c), d)
    x_ref = restate.tdma_batched(a, b, c, d)
    assert np.allclose(x, x_ref, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("name", ["c1_sh_len_so_500x50", "known_sh_vel_so_500x200"])
def test_short_range_and_full_scans_agree(name, monkeypatch):
    """the short-ranged cross-warp inflow (taken when the LU multipliers decay below 1e-30 over a warp) must be
    indistinguishable from the full block-wide scan; plain launches and CUDA-graph replay must agree as well"""
    eng = _engine()
    p = load_golden(name)
    n = 200
    out = {}
    for mode, env in (("default", {}), ("full_scan", {"ION_FULL_SCAN": "1"}), ("no_graphs", {"ION_NO_GRAPHS": "1"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with eng.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"][:n], p["fields"][:n])
            out[mode] = sim.read_g()[0]
        for k in env:
            monkeypatch.delenv(k)
    assert rel_err(out["default"], out["full_scan"]) < 1e-13
    assert rel_err(out["default"], out["no_graphs"]) == 0.0


@pytest.mark.parametrize("name,n", [("sh_vel_so_60x8", None), ("c1_sh_vel_so_500x50", 201), ("known_sh_vel_so_500x200", 130)])
def test_velocity_inter_solve_kernel_matches_pair_local_kernels(name, n, monkeypatch):
    """the out-of-place slab kernel (five operators between two Crank-Nicolson solves in one pass, slab.cuh) against
    the pair-local kernels it replaces; odd step counts exercise the copy back into the home buffer, l_bound = 50 a
    partial last quad, 500 x 200 several slabs and every kind of halo"""
    eng = _engine()
    p = load_golden(name)
    n = len(p["taus"]) if n is None else n
    out = {}
    for mode in ("slab", "pair_local"):
        if mode == "pair_local":
            monkeypatch.setenv("ION_NO_SLAB", "1")
        with eng.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"][:n], p["fields"][:n])
            out[mode] = sim.read_g()[0]
            launches = sim.launch_count
        if mode == "pair_local":
            monkeypatch.delenv("ION_NO_SLAB")
        out[mode + "_launches"] = launches
    assert rel_err(out["slab"], out["pair_local"]) < 1e-13
    assert out["slab_launches"] < out["pair_local_launches"]  # the slab path really ran


def test_velocity_inter_solve_kernel_with_sparse_observations_and_ensemble():
    """observation pattern with fused stretches of odd and even length in between, three ensemble members with
    different pulses: run() with the slab kernel equals member-by-member runs without it"""
    import os

    eng = _engine()
    p = load_golden("sh_vel_so_datastores_120x12")
    n = len(p["taus"])
    fields = np.stack([p["fields"], 0.5 * p["fields"], -1.3 * p["fields"]], axis=1)
    pattern = np.zeros(n, dtype=np.uint8)
    pattern[[2, 3, 9, 14, n - 1]] = 1
    what = eng.nat.OBS_NORM | eng.nat.OBS_INNER_PRODUCTS
    with eng.DeviceSimulation.from_problem(p, batch=3) as sim:
        rec = sim.run(p["taus"], fields, pattern, what)
        g = sim.read_g()
    os.environ["ION_NO_SLAB"] = "1"
    try:
        for b in range(3):
            with eng.DeviceSimulation.from_problem(p) as sim:
                rec_b = sim.run(p["taus"], np.ascontiguousarray(fields[:, b]), pattern, what)
                g_b = sim.read_g()[0]
            assert rel_err(g[b], g_b) < 1e-13
            assert np.max(np.abs(rec[:, b, :] - rec_b[:, 0, :])) < 1e-13
    finally:
        del os.environ["ION_NO_SLAB"]


@pytest.mark.parametrize("name,n", [("sh_len_so_100x10", None), ("c1_sh_len_so_500x50", 201), ("known_sh_len_so_500x200", 130)])
def test_length_gauge_folded_step_matches_two_pass_schedule(name, n, monkeypatch):
    """PROG_LEN_STEP (even sweep folded into the out-of-place odd-pair Crank-Nicolson kernel: one pass per step)
    against the two-pass schedule [ROT even] [ROT-CN-ROT odd]"""
    eng = _engine()
    p = load_golden(name)
    n = len(p["taus"]) if n is None else n
    out = {}
    for mode in ("folded", "two_pass"):
        if mode == "two_pass":
            monkeypatch.setenv("ION_NO_LEN_FOLD", "1")
        with eng.DeviceSimulation.from_problem(p) as sim:
            sim.step(p["taus"][:n], p["fields"][:n])
            out[mode] = sim.read_g()[0]
            out[mode + "_launches"] = sim.launch_count
        if mode == "two_pass":
            monkeypatch.delenv("ION_NO_LEN_FOLD")
    assert rel_err(out["folded"], out["two_pass"]) < 1e-13
    assert out["folded_launches"] < out["two_pass_launches"]


def test_length_gauge_folded_step_with_sparse_observations_and_ensemble():
    import os

    eng = _engine()
    p = load_golden("sh_len_so_datastores_120x12")
    n = len(p["taus"])
    fields = np.stack([p["fields"], 0.5 * p["fields"], -1.3 * p["fields"]], axis=1)
    pattern = np.zeros(n, dtype=np.uint8)
    pattern[[0, 2, 3, 9, 14, n - 1]] = 1
    what = eng.nat.OBS_NORM | eng.nat.OBS_INNER_PRODUCTS
    with eng.DeviceSimulation.from_problem(p, batch=3) as sim:
        rec = sim.run(p["taus"], fields, pattern, what)
        g = sim.read_g()
    os.environ["ION_NO_LEN_FOLD"] = "1"
    try:
        for b in range(3):
            with eng.DeviceSimulation.from_problem(p) as sim:
                rec_b = sim.run(p["taus"], np.ascontiguousarray(fields[:, b]), pattern, what)
                g_b = sim.read_g()[0]
            assert rel_err(g[b], g_b) < 1e-13
            assert np.max(np.abs(rec[:, b, :] - rec_b[:, 0, :])) < 1e-13
    finally:
        del os.environ["ION_NO_LEN_FOLD"]


def test_adi_ensemble_with_different_fields_and_sparse_observations():
    """ION_SH_LEN_ADI (evolution_methods.py:49-77) as a batch: every member has its own field series; observations every 7th
    step go through the CUDA-graph chunks.  Checked against the oracle member by member."""
    from oracle import restate

    eng = _engine()
    p = load_golden("sh_len_adi_64x8")
    n, batch = len(p["taus"]), 3
    scale = np.array([1.0, -0.5, 2.5])
    fields = p["fields"][:, None] * scale[None, :]
    mask = np.zeros(n, dtype=np.uint8)
    mask[6::7] = 1
    with eng.DeviceSimulation.from_problem(p, batch=batch) as sim:
        rec = sim.run(p["taus"], fields, mask, eng.nat.OBS_NORM)
        g = sim.read_g()
    for b in range(batch):
        q = dict(p)
        q["fields"] = fields[:, b]
        ref = restate.run_sh(q, store_every_step=True)
        assert rel_err(g[b], ref["g"]) < TOL
        assert np.max(np.abs(rec[:, b, 0] - ref["norm"][1:][mask.astype(bool)])) < TOL


@pytest.mark.parametrize("L,R", [(70, 150), (513, 40), (9, 4100), (8, 4100), (6, 1000)])
def test_adi_chunk_geometries_against_oracle(L, R):
    """l-pass thread geometries: several chunks per position with PW = 32 / 4 positions per CTA, and an r-segmented mesh."""
    from ionization_b200 import configs, units as u
    from oracle import restate

    eng = _engine()
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=12)
    p["kind"] = "sh_len_adi"
    rng = np.random.default_rng(0)
    g0 = np.array(p["g0"])
    g0 = g0 + 0.05 * (rng.standard_normal(g0.shape) + 1j * rng.standard_normal(g0.shape)) * np.exp(-np.arange(L)[:, None] / 20.0)
    p["g0"] = g0 / np.sqrt(restate.norm(g0, float(p["delta_r"])))
    p["fields"] = np.asarray(p["fields"]) * 30.0 + 1e10  # strong coupling across many channels
    with eng.DeviceSimulation.from_problem(p, with_states=False) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0]
    ref = restate.run_sh(p, store_every_step=False)
    assert rel_err(g, ref["g"]) < TOL


# ---------------------------------------------------------------------------------------------
# edge cases: tiny and ragged meshes, empty calls, step-by-step driving
# ---------------------------------------------------------------------------------------------
def _random_problem(kind, L, R, n_steps, seed=0):
    """a seeded synthetic problem of the fixtures' layout (random H0 with the reference's structure: real diagonal, constant-sign
    off-diagonal), checked against the oracle on the same inputs"""
    rng = np.random.default_rng(seed)
    p = dict(kind=kind, L=L, R=R, r=np.linspace(0.05, 0.05 + 0.1 * (R - 1), R), delta_r=0.1)
    p["h_diag"] = (rng.uniform(0.5, 3.0, (L, R)) + 0.0j)
    p["h_off"] = -rng.uniform(0.4, 0.6, R - 1)
    p["mask"] = np.cos(np.linspace(0, 1.2, R)) ** 0.125
    g0 = rng.standard_normal((L, R)) + 1j * rng.standard_normal((L, R))
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * 0.1)
    p["taus"] = np.full(n_steps, 0.05)
    p["fields"] = rng.uniform(-1, 1, n_steps)
    p["c_l"] = rng.uniform(0.3, 0.6, max(L - 1, 0))
    if kind == "sh_vel_so":
        p["f1_l"] = rng.uniform(0.5, 5.0, max(L - 1, 0))
        p["y_j"] = rng.uniform(0.01, 1.0, R)
        p["z_j"] = rng.uniform(0.01, 0.5, R - 1)
    else:
        p["x_j"] = rng.uniform(0.0, 2.0, R)
    return p


@pytest.mark.parametrize("kind", ["sh_len_so", "sh_vel_so", "sh_len_adi"])
@pytest.mark.parametrize("L,R", [(1, 8), (2, 2), (2, 5), (3, 33), (4, 129), (5, 127), (17, 4), (8, 1025)])
def test_tiny_and_ragged_meshes_against_oracle(kind, L, R):
    from oracle import restate

    eng = _engine()
    p = _random_problem(kind, L, R, 7)
    with eng.DeviceSimulation.from_problem(p, with_states=False) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0]
        norm = sim.observe(eng.nat.OBS_NORM)[0, 0]
    ref = restate.run_sh(p, store_every_step=False)
    assert rel_err(g, ref["g"]) < TOL
    assert abs(norm - ref["norm"][-1]) < TOL * max(1.0, ref["norm"][-1])


@pytest.mark.parametrize("kind", ["sh_len_so", "sh_vel_so", "sh_len_adi"])
def test_empty_call_and_step_by_step_equal_one_call(kind):
    """n_steps = 0 is a no-op; driving the engine one step per call (the callback path of MeshSimulation.run, mesh/sims.py:307)
    gives the same wavefunction as one call (the fused schedules differ, the arithmetic may differ in the last bits)."""
    eng = _engine()
    p = _random_problem(kind, 6, 70, 9, seed=3)
    with eng.DeviceSimulation.from_problem(p, with_states=False) as a, eng.DeviceSimulation.from_problem(p, with_states=False) as b:
        a.step(p["taus"][:0], p["fields"][:0])
        assert rel_err(a.read_g()[0], p["g0"]) < 1e-15
        a.step(p["taus"], p["fields"])
        for n in range(len(p["taus"])):
            b.step(p["taus"][n : n + 1], p["fields"][n : n + 1])
        assert rel_err(a.read_g()[0], b.read_g()[0]) < 1e-13


@pytest.mark.parametrize("R,batch", [(1000, 400), (900, 420)])
def test_ensemble_persistent_prefetch_kernel_matches_one_cta_per_task(monkeypatch, R, batch):
    """csrc/ensemble.cuh: scan ensembles with at least four waves of (pair, member) tasks run the folded length-gauge step in
    persistent CTAs with a cp.async prefetch pipeline.  Same arithmetic as k_unit<LEN_STEP>: compared with that path
    (ION_NO_ENS=1), with sparse observations in between, and member by member with the oracle."""
    from ionization_b200 import configs, units as u
    from oracle import restate

    eng = _engine()
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=8, gauge="LEN", n_steps=14)
    n = len(p["taus"])
    rng = np.random.default_rng(5)
    scale = rng.uniform(-3, 3, batch)
    fields = (np.asarray(p["fields"]) + 2e10)[:, None] * scale[None, :]
    mask = np.zeros(n, dtype=np.uint8)
    mask[4::5] = 1
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ION_NO_ENS", flag)
        with eng.DeviceSimulation.from_problem(p, batch=batch) as sim:
            rec = sim.run(p["taus"], fields, mask, eng.nat.OBS_NORM)
            out[flag] = (sim.read_g(), rec, sim.launch_count)
    assert rel_err(out["0"][0], out["1"][0]) < 1e-13
    assert np.max(np.abs(out["0"][1] - out["1"][1])) < 1e-13
    assert out["0"][2] != out["1"][2]  # the persistent path really ran (it adds a launch for the two single channels)
    for b in (0, 7, batch - 1):
        q = dict(p)
        q["fields"] = fields[:, b]
        ref = restate.run_sh(q, store_every_step=False)
        assert rel_err(out["0"][0][b], ref["g"]) < TOL


def _random_line_problem(kind, Z, n_steps, seed=0):
    rng = np.random.default_rng(seed)
    z = np.linspace(-1.0, 1.0, Z) * 0.05 * Z
    dz = z[1] - z[0]
    p = dict(kind=kind, Z=Z, z=z, delta_z=dz)
    p["h_diag"] = rng.uniform(1.0, 4.0, Z) + 0.0j
    p["h_off"] = -rng.uniform(0.4, 0.6, Z - 1)
    p["w_z"] = z * 0.3
    p["v_pref"] = 0.7
    p["mask"] = np.cos(np.linspace(-1.2, 1.2, Z)) ** 0.125
    g0 = rng.standard_normal(Z) + 1j * rng.standard_normal(Z)
    p["g0"] = g0 / np.sqrt(np.sum(np.abs(g0) ** 2) * dz)
    p["taus"] = np.full(n_steps, 0.05)
    p["fields"] = rng.uniform(-1, 1, n_steps)
    return p


@pytest.mark.parametrize("kind", ["line_len_cn", "line_len_so", "line_vel_so"])
@pytest.mark.parametrize("Z", [2, 3, 5, 33, 128, 129, 1000, 4097])
def test_tiny_and_ragged_line_meshes_against_oracle(kind, Z):
    """LineMesh edge cases (two points, odd sizes, one row past a CTA / a segment boundary) against the oracle"""
    from oracle import restate

    eng = _engine()
    p = _random_line_problem(kind, Z, 6)
    with eng.DeviceSimulation.from_problem(p, with_states=False) as sim:
        sim.step(p["taus"], p["fields"])
        g = sim.read_g()[0, 0]
        norm = sim.observe(eng.nat.OBS_NORM)[0, 0]
    ref = restate.run_line(p, store_every_step=False)
    assert rel_err(g, ref["g"]) < TOL
    assert abs(norm - ref["norm"][-1]) < TOL * max(1.0, ref["norm"][-1])


def test_adi_radial_solve_with_eight_rows_per_thread_equals_the_unit_kernel(monkeypatch):
    """k_adi_r (adi.cuh: eight consecutive rows per thread, LU factors in registers) against k_unit<PROG_CN> (ION_NO_ADI_R=1) and the oracle,
    on a ragged mesh with a mask"""
    from ionization_b200 import configs, units as u
    from oracle import restate

    eng = _engine()
    L, R = 12, 1777
    p = configs.spherical_harmonic_problem(r_bound=0.1 * R * u.bohr_radius, r_points=R, l_bound=L, gauge="LEN", n_steps=10)
    p["kind"] = "sh_len_adi"
    p["mask"] = np.cos(np.linspace(0, 1.1, R)) ** 0.125
    p["fields"] = np.asarray(p["fields"]) * 30.0 + 1e10
    ref = restate.run_sh(p, store_every_step=False)
    out = {}
    for env in ("0", "1"):
        monkeypatch.setenv("ION_NO_ADI_R", env)
        with eng.DeviceSimulation.from_problem(p, with_states=False) as sim:
            sim.step(p["taus"], p["fields"])
            out[env] = sim.read_g()[0]
        assert rel_err(out[env], ref["g"]) < TOL, env
    monkeypatch.delenv("ION_NO_ADI_R")
    assert rel_err(out["0"], out["1"]) < 1e-13
