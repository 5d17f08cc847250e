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
    import dataclasses
    from dumux_b200 import binding as B
    spec = problems.twop_buckleyleverett()
    o = Oracle(spec)
    uo, nso, itso, dtso = o.run_timeloop(spec.initial, T_END, DT0, MAX_DT)
    # (1) step-wise parity: every one of the first time steps, started from the oracle's state, is the same Newton solve on the
    # device -- bit-identical assembly, same iteration count, same shifts, same fields
    u = spec.initial.reshape(-1).copy()
    for k in range(6):
        sp = dataclasses.replace(spec, options=dataclasses.replace(spec.options, dt=float(dtso[k])))
        ok = Oracle(sp)
        e = engine_factory(sp)
        ro, jo = ok.assemble(u, u)
        rg, jg = e.assemble(u.reshape(-1, 2), u.reshape(-1, 2))
        assert np.array_equal(ro, rg) and np.array_equal(jo, jg), k
        un, sto, repo = ok.newton(u, u)
        ug, stg, repg = e.newton(u.reshape(-1, 2), u.reshape(-1, 2))
        assert sto == 0 and stg == 0 and repg.newton_iterations == repo.newton_iterations == itso[k], k
        assert np.linalg.norm(ug - un) <= 1e-12 * np.linalg.norm(un)
        e.close()
        u = un
    # (2) the whole device-resident run.  With zero entry pressure the pressure field is fixed by the fluxes alone and the last
    # Newton update of a step is rounding noise of the size of the shift criterion (1.7e-8 against 4e-9 from start states that
    # differ by 5e-10 Pa), so the Newton counts of a 36-step run -- and with them the step sizes -- are not reproducible between two
    # implementations (or two compilers); the reference's own criterion is what such a run is judged by.
    e = engine_factory(spec)
    ana = _analytic(o)
    for kw in ({}, {"preconditioner": B.PRECOND_AMG}):          # ILU0-BiCGSTAB, and the reference's AMGBiCGSTABIstlSolver (main.cc:78)
        ug, itsg, dtsg = e.run_timeloop(spec.initial, T_END, DT0, MAX_DT, **kw)
        assert abs(len(itsg) - nso) <= 4 and np.isclose(np.sum(dtsg), T_END, rtol=1e-12) and np.max(dtsg) <= MAX_DT * (1 + 1e-12)
        err_com, err_mass = ana.check(spec, ug, T_END)
        assert err_com <= MAX_REL_ERROR and err_mass <= MAX_REL_ERROR, (kw, err_com, err_mass)
        assert np.abs(ug.reshape(-1, 2)[:, 1] - uo.reshape(-1, 2)[:, 1]).max() <= 0.05
