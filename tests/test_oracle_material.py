"""Known-answer tests of the oracle's constitutive relations, mirroring the reference's unit tests:
  test/material/fluidmatrixinteractions/2p/test_material_2p_brookscorey.cc, test_material_2p_vangenuchten.cc and
  testmateriallawfunctions.hh (derivatives vs finite differences, regularised vs raw curves, end points),
  test/common/numericdifferentiation/test_numericdifferentiation.cc (FD formulas through the assembled Jacobian).
"""
import math

import numpy as np
import pytest

from dumux_b200 import problems
from oracle.oracle_py import Oracle, lib

PC, KRW, KRN, DPC, DKRW, DKRN = range(6)


def _law_oracle(law, params, swr=0.1, snr=0.1, regularize=True, reg=()):
    spec = problems.twop_lens((4, 4), law="bc")
    spec.materials = [problems.Material(law, params, swr=swr, snr=snr, regularize=regularize, reg=reg)] * 2
    return Oracle(spec)


def test_det_pow_matches_libm():
    """The shared deterministic pow must be a faithful pow: <= 1 ulp from the correctly rounded result."""
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(1e-6, 1.0, 4000), rng.uniform(1.0, 50.0, 1000), [1.0, 0.5, 1e-12, 0.999999999]])
    y = np.concatenate([rng.uniform(-6.0, 6.0, 5000), [0.5, -0.5, 2.0, 3.0]])
    L = lib()
    worst = 0.0
    for xi, yi in zip(x, y):
        ref = math.pow(xi, yi)
        got = L.orc_pow(xi, yi)
        ulp = math.ulp(ref)
        worst = max(worst, abs(got - ref) / ulp)
    assert worst <= 1.0, worst
    assert L.orc_pow(0.0, 2.5) == 0.0 and L.orc_pow(1.0, -3.3) == 1.0 and L.orc_pow(2.0, 0.0) == 1.0
    assert L.orc_pow(0.0, -0.5) == math.inf


def test_brookscorey_closed_form():
    """test_material_2p_brookscorey.cc parameters: pcEntry 1e4, lambda 2, Swr = Snr = 0.1, pcLowSwe 0.01."""
    o = _law_oracle(problems.LAW_BC, (1e4, 2.0), reg=(0.01,))
    for sw in np.linspace(0.12, 0.88, 31):
        swe = (sw - 0.1) / 0.8
        assert o.law(0, PC, sw) == pytest.approx(1e4 * swe ** -0.5, rel=1e-14)
        assert o.law(0, KRW, sw) == pytest.approx(swe ** 4.0, rel=1e-13)
        assert o.law(0, KRN, sw) == pytest.approx((1 - swe) ** 2 * (1 - swe ** 2.0), rel=1e-13)
    # end points (the reference checks pc at the end points to 1e-10): pc(Swe=1) = pcEntry, kr end points
    assert o.law(0, PC, 0.9) == pytest.approx(1e4, abs=1e-10 * 1e4)
    assert o.law(0, KRW, 0.9) == 1.0 and o.law(0, KRN, 0.9) == 0.0
    assert o.law(0, KRW, 0.1) == 0.0 and o.law(0, KRN, 0.1) == 1.0
    # regularisation: linear continuation below pcLowSwe and above Swe = 1, continuous at the thresholds
    lo_sw = 0.1 + 0.8 * 0.01
    assert o.law(0, PC, lo_sw - 1e-9) == pytest.approx(o.law(0, PC, lo_sw + 1e-9), rel=1e-6)
    slope = (o.law(0, PC, 0.02) - o.law(0, PC, 0.0)) / 0.02
    assert slope == pytest.approx(o.law(0, DPC, 0.05), rel=1e-9)           # straight line below the threshold
    assert o.law(0, PC, 0.95) < 1e4 and o.law(0, PC, 0.95) == pytest.approx(
        1e4 + (0.95 - 0.9) / 0.8 * (-1e4 / 2.0), rel=1e-12)


@pytest.mark.parametrize("law,params,reg", [(problems.LAW_BC, (1e4, 2.0), (0.01,)),
                                            (problems.LAW_VG, (6.66e-5, 3.652, 0.5), (0.01, 0.99, 0.1, 0.9)),
                                            (problems.LAW_VG, (0.0037, 4.7, 0.5), (0.01, 0.99, 0.1, 0.9))])
@pytest.mark.parametrize("regularize", [True, False])
def test_derivatives_match_finite_differences(law, params, reg, regularize):
    """testmateriallawfunctions.hh: every d/dSw against a central difference of the curve itself."""
    o = _law_oracle(law, params, regularize=regularize, reg=reg)
    lo, hi = (0.02, 0.98) if regularize else (0.15, 0.85)
    for sw in np.linspace(lo, hi, 41):
        h = 1e-6
        for f, df in ((PC, DPC), (KRW, DKRW), (KRN, DKRN)):
            num = (o.law(0, f, sw + h) - o.law(0, f, sw - h)) / (2 * h)
            ana = o.law(0, df, sw)
            scale = max(abs(ana), abs(num), 1e-6 if f != PC else 1.0)
            assert abs(num - ana) <= 2e-5 * scale + 1e-7, (f, sw, num, ana)


def test_vangenuchten_regularised_equals_raw_inside_thresholds():
    """test_material_2p_vangenuchten.cc: the regularised law equals the raw law between the thresholds, is monotone and
    continuous across them."""
    reg = _law_oracle(problems.LAW_VG, (6.66e-5, 3.652, 0.5), regularize=True, reg=(0.01, 0.99, 0.1, 0.9))
    raw = _law_oracle(problems.LAW_VG, (6.66e-5, 3.652, 0.5), regularize=False)
    for swe in np.linspace(0.11, 0.89, 40):
        sw = 0.1 + 0.8 * swe
        for f in (PC, KRW, KRN):
            assert reg.law(0, f, sw) == raw.law(0, f, sw)
    sws = np.linspace(0.0, 1.0, 2001)
    pc = np.array([reg.law(0, PC, s) for s in sws])
    krw = np.array([reg.law(0, KRW, s) for s in sws])
    krn = np.array([reg.law(0, KRN, s) for s in sws])
    assert np.all(np.diff(pc) <= 1e-9) and np.all(np.diff(krw) >= -1e-12) and np.all(np.diff(krn) <= 1e-12)
    assert np.abs(np.diff(pc)).max() < 0.02 * pc.max()          # no jump at any threshold
    assert np.abs(np.diff(krw)).max() < 5e-3 and np.abs(np.diff(krn)).max() < 5e-3
    assert reg.law(0, PC, 0.9) == pytest.approx(0.0, abs=1e-10)   # pc(Swe = 1) = 0 for van Genuchten


@pytest.mark.parametrize("method,order", [(1, 1), (-1, 1), (0, 2), (5, 4)])
def test_numeric_differentiation_methods(method, order):
    """test_numericdifferentiation.cc: forward/backward/central/5-point quotients against analytic derivatives.  The
    storage term phi*rho_n*S_n*V/dt is linear in S_n and the flux term smooth, so the diagonal entry d r_n / d S_n of an
    isolated cell (K -> 0: no fluxes) must equal phi*rho_n*V/dt for every method."""
    spec = problems.twop_lens((3, 3), law="bc")
    spec.K = np.full_like(spec.K, 1e-30)      # (K = 0 would give 0/0 in the gravity term, in DuMux too)
    spec.options.fd_method = method
    spec.options.base_eps = 1e-6
    o = Oracle(spec)
    cur = spec.initial.copy()
    cur[:, 1] = 0.3
    res, jac = o.assemble(cur, spec.initial)
    vol = (6.0 / 3) * (4.0 / 3)
    expect_n = 0.4 * 1460.0 * vol / 250.0
    expect_w = -0.4 * 1000.0 * vol / 250.0
    for I in range(9):
        kd = [k for k in range(o.rowptr[I], o.rowptr[I + 1]) if o.colidx[k] == I][0]
        blk = jac[kd * 4:(kd + 1) * 4].reshape(2, 2)
        assert blk[1, 1] == pytest.approx(expect_n, rel=1e-8)
        assert blk[0, 1] == pytest.approx(expect_w, rel=1e-8)
        assert blk[0, 0] == 0.0 and blk[1, 0] == 0.0            # incompressible: no pressure dependence of the storage


def test_fd_epsilon_rule():
    """numericepsilon.hh:47-51 / numericdifferentiation.hh:36-41: eps = baseEps*magnitude if set else baseEps*(|x|+1):
    with a linear residual the quotient is exact, so only the pattern of the Jacobian may depend on eps."""
    a = problems.onep_incompressible((6, 6), numdiff_params=True)
    o = Oracle(a)
    x = np.full(36, 1.5e5)
    _, j1 = o.assemble(x)
    a.options.base_eps, a.options.privar_magnitude = 1e-3, (-1.0, -1.0)
    _, j2 = Oracle(a).assemble(x)
    assert np.abs(j1 - j2).max() <= 1e-9 * np.abs(j1).max()
