"""INTEGRATION.md sections 1-2 as executable code: the reference's OWN objects driven through the C-ABI.
`reference`: needs /root/reference (this container); `gpu`: needs a CUDA device (the box, where the reference is absent) -- the
end-to-end test therefore runs only where both exist and skips cleanly otherwise; the input extraction is checked here on the CPU."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import refenv

pytestmark = pytest.mark.reference
needs_reference = pytest.mark.skipif(not refenv.reference_available(), reason="the reference tree is only mounted in the build container")


def _reference_sim(ion, u, method=None, gauge="LEN", R=100, L=10, steps=60):
    rb = 30 * u.bohr_radius
    ops = ion.mesh.SphericalHarmonicLengthGaugeOperators() if gauge == "LEN" else ion.mesh.SphericalHarmonicVelocityGaugeOperators()
    return ion.mesh.SphericalHarmonicSpecification(
        "binding", r_bound=rb, r_points=R, l_bound=L, time_initial=-steps / 2 * u.asec, time_final=steps / 2 * u.asec, time_step=1 * u.asec,
        electric_potential=ion.potentials.SincPulse(pulse_width=20 * u.asec, fluence=1 * u.Jcm2, phase=0), use_numeric_eigenstates=False,
        test_states=[ion.states.HydrogenBoundState(n, l) for n in range(1, 4) for l in range(n)],
        mask=ion.potentials.RadialCosineMask(inner_radius=0.8 * rb, outer_radius=rb, smoothness=8), operators=ops,
        evolution_method=method if method is not None else ion.mesh.SplitInteractionOperator(), store_data_every=1,
    ).to_sim()


@needs_reference
@pytest.mark.parametrize("gauge, fixture, R, L, steps", [("LEN", "sh_len_so_100x10", 100, 10, 60), ("VEL", "sh_vel_so_60x8", 60, 8, 40)])
def test_inputs_extracted_from_reference_objects_equal_the_fixture(gauge, fixture, R, L, steps):
    from ionization_b200 import reference_binding as rb

    ion = refenv.import_reference()
    import simulacra.units as u

    ref = load_golden(fixture)
    p = rb.extract_problem(_reference_sim(ion, u, gauge=gauge, R=R, L=L, steps=steps), u)
    assert p["kind"] == str(ref["kind"])
    for key in ("h_diag", "h_off", "c_l", "mask", "g0", "state_rows") + (("x_j",) if gauge == "LEN" else ("f1_l", "y_j", "z_j")):
        assert rel_err(p[key], ref[key]) < 1e-14, key
    assert np.array_equal(p["state_l"], ref["state_l"])


@needs_reference
@pytest.mark.gpu
def test_reference_run_loop_with_the_device_evolution_method_and_tdma():
    from ionization_b200 import reference_binding as rb

    ion = refenv.import_reference()
    import simulacra.units as u

    cpu = _reference_sim(ion, u)
    cpu.run()
    dev = _reference_sim(ion, u, method=rb.make_evolution_method(ion)())
    dev.run()  # the reference's own MeshSimulation.run(); every evolve() is one ion_sim_step
    assert rel_err(dev.mesh.g, cpu.mesh.g) < 1e-10
    assert np.max(np.abs(dev.data.norm - cpu.data.norm)) < 1e-10
    original = ion.cy.tdma
    try:
        rb.bind_tdma(ion)
        tdma = _reference_sim(ion, u)
        tdma.run()  # the reference's operators, its Thomas solve replaced by ion_tdma_c128
    finally:
        ion.cy.tdma = original
    assert rel_err(tdma.mesh.g, cpu.mesh.g) < 1e-10
