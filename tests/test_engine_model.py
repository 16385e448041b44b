"""The ALGORITHM of the CUDA engine (tests/engine_model.py: chunked affine-scan Crank-Nicolson with x = 2 M^-1 g - g,
pair-local kernel schedule, cross-step fusion of the even rotations), checked on the CPU against reference fixtures."""
import pytest

from conftest import load_golden, rel_err
from engine_model import run_sh_model


@pytest.mark.parametrize("name", ["sh_len_so_100x10", "sh_len_so_101x10", "sh_vel_so_60x8", "sh_vel_so_61x8", "sh_vel_so_datastores_120x12"])
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("M", [2, 4, 8])
def test_engine_algorithm_matches_reference(name, fused, M):
    p = load_golden(name)
    assert rel_err(run_sh_model(p, M=M, fused=fused), p["g_final"]) < 1e-12


@pytest.mark.parametrize("name", ["sh_len_adi_64x8", "sh_len_adi_65x9"])
@pytest.mark.parametrize("CL", [2, 8])
def test_adi_l_pass_algorithm_matches_reference(name, CL):
    """ION_SH_LEN_ADI (csrc/adi.cuh): real pivots through composed Moebius maps + chunked affine recurrences."""
    from engine_model import run_sh_adi_model

    p = load_golden(name)
    assert rel_err(run_sh_adi_model(p, CL=CL), p["g_final"]) < 1e-12
