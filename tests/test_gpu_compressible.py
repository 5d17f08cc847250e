"""GPU parity for the compressible 1p configuration (BASELINE config 2: tabulated H2O, log-normal permeability, 3-D):
assembly bit-identical to the oracle, Newton iteration counts and fields over the reference's check-point time loop."""
import os

import numpy as np
import pytest

from dumux_b200 import problems, timeloop
from oracle.oracle_py import Oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("cells,lognormal", [((10, 10), False), ((37, 11, 19), True)])
@pytest.mark.parametrize("method", [1, 0])
def test_assembly_tabulated_fluid(engine_factory, cells, lognormal, method):
    spec = problems.onep_compressible(cells, lognormal=lognormal)
    spec.options.fd_method = method
    rng = np.random.RandomState(5)
    prev = spec.initial + rng.uniform(-2e4, 5e4, size=spec.initial.shape)
    cur = spec.initial + rng.uniform(-2e4, 5e4, size=spec.initial.shape)
    o = Oracle(spec)
    res_o, jac_o = o.assemble(cur, prev)
    e = engine_factory(spec)
    res_g, jac_g = e.assemble(cur, prev)
    assert np.array_equal(res_g, res_o)
    assert np.abs(jac_g - jac_o).max() <= 1e-10 * np.abs(jac_o).max()
    if method == 1:
        assert np.array_equal(jac_g, jac_o)


def test_pressure_outside_table_is_reported(engine_factory):
    """TabulatedComponent returns NaN outside its temperature range and extrapolates linearly in p; a NaN state must surface
    as a non-finite residual status, not as silent garbage."""
    from dumux_b200.binding import DmxError
    spec = problems.onep_compressible((6, 6))
    cur = spec.initial.copy()
    cur[7, 0] = np.inf
    e = engine_factory(spec)
    with pytest.raises(DmxError):
        e.assemble(cur, spec.initial)


def test_instationary_checkpoint_loop_matches_oracle_and_golden(engine_factory):
    """The reference main's loop (CheckPointTimeLoop, periodic check points tEnd/10) with MaxTimeStepSize = tEnd/10, so that
    the step sequence does not depend on the Newton counts: on this nearly linear problem the last Newton shift sits within
    a small factor of MaxRelativeShift = 1e-8 (FD-Jacobian noise floor), so the COUNT may differ by one between two
    correct implementations whose dot products round differently (sequential sum in the oracle, tree on the GPU)."""
    spec = problems.onep_compressible((10, 10))

    def mk():
        loop = timeloop.CheckPointTimeLoop(0.0, 0.01, 0.1)
        loop.set_max_time_step_size(0.01)
        loop.set_periodic_check_point(0.01)
        return loop

    uo, its_o, dts_o = Oracle(spec).run_instationary(spec.initial, mk())
    e = engine_factory(spec)
    ug, its_g, dts_g = e.run_instationary(spec.initial, mk())
    assert np.array_equal(dts_g, dts_o) and len(dts_g) == 10
    assert max(abs(a - b) for a, b in zip(its_g, its_o)) <= 1
    assert np.linalg.norm(ug - uo) <= 1e-8 * np.linalg.norm(uo)
    g = np.load(os.path.join(GOLDEN, "test_1p_cc.npz"))["p"].astype(np.float64)
    assert np.abs(ug.ravel() / g - 1).max() < 1e-4


def test_stationary_newton_matches_oracle_and_golden(engine_factory):
    """test_1p_compressible_stationary_tpfa: one stationary Newton solve, compared by the reference with test_1p_cc-reference.vtu"""
    import dataclasses
    spec = problems.onep_compressible((10, 10))
    spec.options = dataclasses.replace(spec.options, stationary=True)
    uo, sto, repo = Oracle(spec).newton(spec.initial, spec.initial)
    e = engine_factory(spec)
    ug, stg, repg = e.newton(spec.initial, spec.initial)
    assert sto == 0 and stg == 0 and repg.newton_iterations == repo.newton_iterations
    assert np.linalg.norm(ug - uo) <= 1e-8 * np.linalg.norm(uo)
    g = np.load(os.path.join(GOLDEN, "test_1p_cc.npz"))["p"].astype(np.float64)
    d = np.abs(ug - g)
    assert np.all(d <= 1e-2 * np.maximum(np.abs(ug), np.abs(g)))


def test_newton_step_3d_lognormal(engine_factory):
    spec = problems.onep_compressible((24, 16, 20), lognormal=True, dt=0.002)
    uo, sto, repo = Oracle(spec).newton(spec.initial, spec.initial)
    e = engine_factory(spec)
    ug, stg, repg = e.newton(spec.initial, spec.initial)
    assert sto == 0 and stg == 0 and repg.newton_iterations == repo.newton_iterations
    assert np.linalg.norm(ug - uo) <= 1e-8 * np.linalg.norm(uo)
