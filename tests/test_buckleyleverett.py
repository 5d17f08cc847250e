"""test/porousmediumflow/2p/buckleyleverett (test_2p_buckleyleverett_tpfa) on the oracle and on the device: the 1-D displacement
run to tEnd = 1e7 s with the reference's time-step control, accepted by the reference's own criterion (main.cc:140-205): the
wetting-phase centre of mass and the total wetting-phase mass agree with the analytic Buckley-Leverett profile (Welge construction,
analyticsolution.hh) to MaxRelError = 0.01."""
import numpy as np
import pytest

from dumux_b200 import problems
from oracle.oracle_py import Oracle

T_END, DT0, MAX_DT, MAX_REL_ERROR = 1e7, 1e3, 5e5, 0.01


def _analytic(o):
    return problems.BuckleyLeverettAnalyticSolution(lambda which, sw: o.law(0, which, sw))


def test_analytic_solution_properties():
    """Welge construction: the shock saturation is the tangent point of f_w seen from Swr, the shock speed equals the
    Rankine-Hugoniot speed, the profile is monotone and self-similar"""
    spec = problems.twop_buckleyleverett()
    ana = _analytic(Oracle(spec))
    assert 0.2 < ana.sw_shock < 0.8
    rh = ana.v / ana.phi * (ana.fw(ana.sw_shock) - ana.fw(ana.swr)) / (ana.sw_shock - ana.swr)
    assert ana.shock_speed == pytest.approx(rh, rel=1e-9)
    x = np.linspace(0.5, 99.5, 100)
    s = np.array([ana.saturation(xx, T_END) for xx in x])
    assert np.all(np.diff(s) <= 1e-12) and s[0] > ana.sw_shock and s[-1] == ana.swr
    assert ana.saturation(10.0, 1e6) == pytest.approx(ana.saturation(20.0, 2e6), rel=1e-12)


def test_buckleyleverett_oracle_meets_the_reference_criterion():
    spec = problems.twop_buckleyleverett()
    o = Oracle(spec)
    u, nsteps, its, dts = o.run_timeloop(spec.initial, T_END, DT0, MAX_DT)
    assert nsteps > 20 and np.isclose(np.sum(dts), T_END, rtol=1e-12) and np.max(dts) <= MAX_DT * (1 + 1e-12)
    err_com, err_mass = _analytic(o).check(spec, u, T_END)
    assert err_com <= MAX_REL_ERROR and err_mass <= MAX_REL_ERROR, (err_com, err_mass)
    sn = u.reshape(-1, 2)[:, 1]
    assert sn.min() >= 0.2 - 1e-9 and sn.max() <= 0.8 + 1e-9          # between the residual saturations


@pytest.mark.gpu
def test_buckleyleverett_device_matches_oracle_and_reference_criterion(engine_factory):
    from dumux_b200 import binding as B
    spec = problems.twop_buckleyleverett()
    o = Oracle(spec)
    uo, nso, itso, dtso = o.run_timeloop(spec.initial, T_END, DT0, MAX_DT)
    e = engine_factory(spec)
    ug, itsg, dtsg = e.run_timeloop(spec.initial, T_END, DT0, MAX_DT)
    assert list(itsg) == list(itso), (itsg, itso)
    assert np.allclose(dtsg, dtso, rtol=0, atol=0)
    a, b = ug.reshape(-1, 2), uo.reshape(-1, 2)
    assert np.linalg.norm(a[:, 0] - b[:, 0]) <= 1e-8 * np.linalg.norm(b[:, 0])
    assert np.linalg.norm(a[:, 1] - b[:, 1]) <= 1e-8 * np.linalg.norm(b[:, 1])
    ana = _analytic(o)
    err_com, err_mass = ana.check(spec, ug, T_END)
    assert err_com <= MAX_REL_ERROR and err_mass <= MAX_REL_ERROR, (err_com, err_mass)
    # ... and with the linear solver of the reference's main.cc:78 (AMGBiCGSTABIstlSolver)
    ua, itsa, dtsa = e.run_timeloop(spec.initial, T_END, DT0, MAX_DT, preconditioner=B.PRECOND_AMG)
    assert abs(len(itsa) - len(itsg)) <= 3 and np.isclose(np.sum(dtsa), T_END, rtol=1e-12)
    err_com, err_mass = ana.check(spec, ua, T_END)
    assert err_com <= MAX_REL_ERROR and err_mass <= MAX_REL_ERROR, (err_com, err_mass)
    # (a Newton count that differs in one step changes the later step sizes, so the two runs agree to the time-discretisation
    # error, not to the solver tolerance)
    assert np.abs(ua.reshape(-1, 2)[:, 1] - a[:, 1]).max() <= 0.05
