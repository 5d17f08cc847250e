"""Host time-step control (dumux_b200/timeloop.py) against the rules of dumux/common/timeloop.hh, and the compressible 1p
configuration (BASELINE config 2) of the oracle against the reference's golden field."""
import os

import numpy as np

from dumux_b200 import iapws, problems, timeloop

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_timeloop_clamps_to_end_and_finishes():
    loop = timeloop.TimeLoop(0.0, 250.0, 1000.0)
    dts = []
    while not loop.finished():
        dts.append(loop.dt)
        loop.advance_time_step()
        loop.set_time_step_size(loop.dt * 1.5)
    assert abs(sum(dts) - 1000.0) < 1e-9 and dts[0] == 250.0 and dts[-1] <= 1000.0 - sum(dts[:-1]) + 1e-12
    assert loop.max_time_step_size() == 0.0


def test_checkpoint_timeloop_hits_every_periodic_check_point():
    loop = timeloop.CheckPointTimeLoop(0.0, 0.002, 0.1)
    loop.set_periodic_check_point(0.01)
    times, flags = [], []
    while not loop.finished():
        loop.advance_time_step()
        times.append(loop.time)
        flags.append(loop.is_check_point)
        loop.set_time_step_size(timeloop.suggest_time_step_size(loop.dt, 4))
    hit = [t for t, f in zip(times, flags) if f]
    assert np.allclose(hit, np.arange(1, 11) * 0.01, rtol=0, atol=1e-12)
    assert abs(times[-1] - 0.1) < 1e-12
    # a step that would leave less than 20 % of itself to the check point is stretched onto it (timeloop.hh:549-560)
    dts = np.diff([0.0] + times)
    assert dts.min() > 1e-4


def test_suggest_time_step_size():
    assert timeloop.suggest_time_step_size(1.0, 10) == 1.0
    assert abs(timeloop.suggest_time_step_size(1.0, 15) - 1.0 / 1.5) < 1e-15
    assert abs(timeloop.suggest_time_step_size(1.0, 4) - (1.0 + 0.6 / 1.2)) < 1e-15


def test_iapws_if97_verification_values():
    """IAPWS-IF97 Table 5 (region 1): v(300 K, 3 MPa) = 0.100215168e-2 m^3/kg; Table 35: p_sat(300 K) = 0.353658941e-2 MPa;
    IAPWS 2008 viscosity check value mu(298.15 K, 998 kg/m^3) = 889.735100e-6 Pa s."""
    assert abs(iapws.volume_region1(300.0, 3e6) / 0.100215168e-2 - 1) < 1e-6   # DuMux uses Rs = 8.314472/18.01518e-3, IF97 0.461526 kJ/(kg K)
    assert abs(iapws.saturation_pressure(300.0) / 0.353658941e4 - 1) < 1e-8
    assert abs(iapws.viscosity(298.15, 998.0) / 889.735100e-6 - 1) < 1e-6
    t = iapws.tabulated_h2o()
    assert t["rho"].shape == (2000,) and np.all(t["pmin"] == 1e4) and np.all(t["pmax"] == 1e6)
    assert 995.0 < t["rho"].min() and t["rho"].max() < 1001.0 and 9e-4 < t["mu"].min() and t["mu"].max() < 1.9e-3


def test_oracle_1p_compressible_instationary_matches_golden():
    """test_1p_compressible_instationary_tpfa: CheckPointTimeLoop(0, 0.002, 0.1), periodic check points tEnd/10
    (main.cc:118-150), compared with test/references/test_1p_cc-reference.vtu (the reference's own regression file for
    this test) at the reference's fuzzy tolerance."""
    from oracle.oracle_py import Oracle
    spec = problems.onep_compressible((10, 10))
    o = Oracle(spec)
    loop = timeloop.CheckPointTimeLoop(0.0, 0.002, 0.1)
    loop.set_periodic_check_point(0.01)
    u, its, dts = o.run_instationary(spec.initial, loop)
    g = np.load(os.path.join(GOLDEN, "test_1p_cc.npz"))["p"].astype(np.float64)
    assert np.abs(u.ravel() / g - 1).max() < 1e-4          # reference bar: 1e-2 relative
    assert abs(sum(dts) - 0.1) < 1e-12 and max(its) <= 6 and min(its) >= 2
