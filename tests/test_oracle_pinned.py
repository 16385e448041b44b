"""The oracle (oracle/restate.py numpy, oracle/c/restate.c C) pinned to the UNMODIFIED reference:
fixtures produced by running /root/reference under the shim (oracle/make_golden.py), the reference's own known
answers (dev/meshes/mesh_refactoring_helper.py:204-251) and, when present, the compiled cy.tdma (oracle/_ref)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_err
from oracle import cport, restate

SMALL = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(f).startswith(("known_", "c1_")))
KNOWN = {  # dev/meshes/mesh_refactoring_helper.py:204-251
    "known_sh_len_adi_500x200": 0.312910470190,
    "known_sh_len_so_500x200": 0.312928752359,
    "known_sh_vel_so_500x200": 0.319513371899,
    "known_line_len_cn_4096": 0.370010185740,
    "known_line_len_so_4096": 0.370008474418,
    "known_line_vel_so_4096": 0.370924310122,
}


@pytest.mark.parametrize("name", SMALL)
def test_numpy_restatement_matches_reference_run(name):
    p = load_golden(name)
    if str(p["kind"]).startswith("sh"):
        out = restate.run_sh(p, state_l=p["state_l"], state_rows=p["state_rows"])
    else:
        out = restate.run_line(p, state_rows=p["state_rows"])
    assert rel_err(out["g"], p["g_final"]) < 1e-13
    # the restatement records every time index; fixtures with store_data_every > 1 hold the data times only
    idx = np.searchsorted(p["times"], p["data_times"]) if len(p["norm"]) != len(out["norm"]) else slice(None)
    assert np.max(np.abs(out["norm"][idx] - p["norm"])) < 1e-13
    assert np.max(np.abs(out["inner_products"][idx] - p["inner_products"])) < 1e-13


@pytest.mark.parametrize("name", [n for n in SMALL if "adi" not in n] + ["c1_sh_len_so_500x50", "c1_sh_vel_so_500x50"])
def test_c_restatement_matches_reference_run(name):
    p = load_golden(name)
    g = cport.sh_steps(p) if str(p["kind"]).startswith("sh") else cport.line_steps(p)
    assert rel_err(g, p["g_final"]) < 1e-12


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_reference_known_answers(name):
    """the fixture was produced by the reference itself; its final overlap is the reference's published known answer
    to the 12 printed digits (one unit in the last printed digit allowed), and the oracle reproduces the run"""
    p = load_golden(name)
    assert abs(float(p["initial_state_overlap_final"]) - KNOWN[name]) < 2e-12
    kind = str(p["kind"])
    if kind == "sh_len_adi":
        return  # 800 steps of the python-loop ADI restatement take minutes; covered on the small fixtures
    g = cport.sh_steps(p) if kind.startswith("sh") else cport.line_steps(p)
    assert rel_err(g, p["g_final"]) < 1e-11
    i0 = int(p["initial_state_index"])
    if kind.startswith("sh"):
        ip = restate.inner_product_rows(g, p["state_l"][i0 : i0 + 1], p["state_rows"][i0 : i0 + 1], float(p["delta_r"]))[0]
    else:
        ip = restate.inner_product_full(p["state_rows"][i0], g, float(p["delta_z"]))
    assert abs(abs(ip) ** 2 - KNOWN[name]) < 2e-12


@pytest.mark.parametrize("name", ["sh_len_so_datastores_120x12", "sh_vel_so_datastores_120x12"])
def test_observables_restatement(name):
    p = load_golden(name)
    g = p["g_final"]
    ipm = float(p["delta_r"])
    assert abs(restate.norm(g, ipm) - p["norm"][-1]) < 1e-13
    assert rel_err(restate.norm_by_l(g, ipm), p["norm_by_l"][-1]) < 1e-12
    assert abs(restate.r_expectation(g, p["r"], ipm) - p["r_expectation"][-1]) < 1e-12 * abs(p["r_expectation"][-1])
    assert abs(restate.sh_z_expectation(g, p["c_l"], p["r"], ipm) - p["z_expectation"][-1]) < 1e-12 * abs(p["r_expectation"][-1])
    assert abs(restate.h0_expectation(g, p["h_diag"], p["h_off"], ipm) - p["internal_energy"][-1]) < 1e-12 * abs(p["internal_energy"][-1])
    for k, rad in enumerate(p["norm_within_radii"]):
        assert abs(restate.norm_within_radius(g, p["r"], rad, ipm) - p["norm_within_radius"][-1, k]) < 1e-13
    if "total_energy" in p:
        e = p["efield_half_at_data_times"][-1]
        tot = restate.sh_len_total_energy_expectation(g, p["h_diag"], p["h_off"], p["c_l"], p["x_j"], e, ipm)
        assert abs(tot - p["total_energy"][-1]) < 1e-12 * abs(p["total_energy"][-1])


# ---- tdma: tests/test_tdma.py:12-26 of the reference, seeded -----------------------------------------------------
@pytest.mark.parametrize("n", [2, 3, 10, 257, 1000])
def test_tdma_restatements_agree_with_dense_inverse(n):
    from scipy import sparse

    rng = np.random.default_rng(n)
    crs = lambda k: rng.random(k) + 1j * rng.random(k)  # tests/conftest.py:4-5
    a, b, c, d = crs(n - 1), crs(n), crs(n - 1), crs(n)
    dia = sparse.diags([a, b, c], offsets=[-1, 0, 1])
    inv_x = np.linalg.inv(dia.toarray()).dot(d)
    for x in (restate.tdma(a, b, c, d), cport.tdma(a, b, c, d), restate.tdma_batched(a[None], b[None], c[None], d[None])[0]):
        assert np.allclose(x, inv_x)
        assert np.allclose(dia.dot(x), d)


def _load_ref_cy():
    paths = glob.glob(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "cy*.so"))
    if not paths:
        return None
    spec = importlib.util.spec_from_file_location("cy", paths[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tdma_restatement_equals_the_compiled_reference_cy_tdma():
    """oracle/_ref/cy*.so is the reference's own cy.pyx compiled from /root/reference (make -C oracle ref)"""
    cy = _load_ref_cy()
    if cy is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    from scipy import sparse

    rng = np.random.default_rng(7)
    for n in (2, 5, 100, 999):
        crs = lambda k: rng.random(k) + 1j * rng.random(k)
        a, b, c, d = crs(n - 1), crs(n) + 1.5, crs(n - 1), crs(n)
        dia = sparse.diags([a, b, c], offsets=[-1, 0, 1]).todia()
        x_ref = cy.tdma(dia, d)
        assert np.max(np.abs(restate.tdma(a, b, c, d) - x_ref)) <= 1e-14 * np.max(np.abs(x_ref))
        assert np.max(np.abs(cport.tdma(a, b, c, d) - x_ref)) <= 1e-14 * np.max(np.abs(x_ref))
