"""Pins the CPU oracle to the reference's own golden fields.

The fixtures under tests/golden/ were extracted from /root/reference/test/references/*.vtu by
tests/golden/make_golden.py (Float32 cell data).  The reference's own regression bar is the fuzzy comparison of
bin/testing/dumux_runtest.py / fuzzycomparevtu.py: |a-b| <= 1.5e-7 or |a-b| <= 1e-2*max(|a|,|b|); the oracle is held to
that bar and, where Float32 storage allows, to a much tighter one (noted per test).
"""
import os

import numpy as np
import pytest

from dumux_b200 import problems
from oracle.oracle_py import Oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _fuzzy_ok(a, ref, rel=1e-2, abs_=1.5e-7):
    d = np.abs(a - ref)
    return np.all((d <= abs_) | (d <= rel * np.maximum(np.abs(a), np.abs(ref))))


def _one_linear_step(spec, reduction=1e-13):
    """test/porousmediumflow/1p/incompressible/main.cc:150-161: assemble, solve J dx = r, x -= dx."""
    o = Oracle(spec)
    x = spec.initial.reshape(-1).copy()
    res, jac = o.assemble(x)
    dx, st, its, red = o.solve(jac, res, reduction=reduction)
    assert st == 0
    return x - dx, o, its


def test_1p_incompressible_10x10_reference_vtu():
    """test_1p_incompressible_tpfa_numdiff (BaseEpsilon 0.1, PriVarMagnitude 1e5) -> test_1p_cc-reference.vtu"""
    x, o, its = _one_linear_step(problems.onep_incompressible((10, 10)))
    g = np.load(os.path.join(GOLDEN, "test_1p_cc.npz"))["p"].astype(np.float64)
    assert _fuzzy_ok(x, g)
    # the VTU stores 6 significant decimal digits: agreement to 5e-6 relative is the best possible
    assert np.abs(x / g - 1).max() < 5e-6
    # the linear problem is solved exactly by one Newton step: the residual at x vanishes
    r2, _ = o.assemble(x, jacobian=False)
    r0, _ = o.assemble(np.zeros_like(x), jacobian=False)
    assert np.linalg.norm(r2) <= 1e-9 * np.linalg.norm(r0)


def test_1p_compressible_stationary_reference_vtu():
    """test_1p_compressible_stationary_tpfa (test/porousmediumflow/1p/compressible/stationary: the compressible 1p problem with
    tabulated water solved as ONE stationary Newton solve with ILUBiCGSTABIstlSolver, main.cc:100-114) is compared by the reference
    with the SAME test_1p_cc-reference.vtu as the incompressible test, at the fuzzy bar (relative 1e-2)."""
    import dataclasses
    spec = problems.onep_compressible((10, 10))
    spec.options = dataclasses.replace(spec.options, stationary=True)
    u, st, rep = Oracle(spec).newton(spec.initial, spec.initial)
    assert st == 0 and 2 <= rep.newton_iterations <= 5
    g = np.load(os.path.join(GOLDEN, "test_1p_cc.npz"))["p"].astype(np.float64)
    assert _fuzzy_ok(u, g)
    assert np.abs(u / g - 1).max() < 1e-4          # the water's compressibility moves the pressure by 2e-5 of its value


def test_1p_isothermal_tpfa_reference_vtu():
    """test_1p_tpfa (test/porousmediumflow/1p/isothermal: SimpleH2O, one implicit Euler step dt = 1 s with the Newton solver,
    Dirichlet p = 1e5 (2 - y) at top and bottom, gravity, lens) -> test_1p_cc-reference.vtu"""
    import dataclasses
    spec = problems.onep_compressible((10, 10))
    spec = dataclasses.replace(spec, fluid_table=None, options=dataclasses.replace(spec.options, dt=1.0))
    u, st, rep = Oracle(spec).newton(spec.initial, spec.initial)
    assert st == 0
    g = np.load(os.path.join(GOLDEN, "test_1p_cc.npz"))["p"].astype(np.float64)
    assert _fuzzy_ok(u, g) and np.abs(u / g - 1).max() < 5e-6


def test_1p_incompressible_tpfa_extrude_constant_velocity():
    """test_1p_incompressible_tpfa_extrude (-Problem.ExtrusionFactor 10 -Problem.CheckIsConstantVelocity true -Problem.EnableGravity
    false): homogeneous K, the analytic Jacobian, extrusion factor 10 in transmissibilities and fluxes"""
    import dataclasses
    spec = problems.onep_extrude()
    o = Oracle(spec)
    u, st, rep = o.newton(spec.initial, spec.initial)
    assert st == 0
    dev_y, dev_x = problems.constant_velocity_check(spec, o.volume_flux(u))
    assert dev_y <= 1e-8 and dev_x <= 1e-10
    # the pressure does not depend on the extrusion factor
    s1 = problems.onep_extrude(extrusion=1.0)
    u1, st1, _ = Oracle(s1).newton(s1.initial, s1.initial)
    assert np.abs(u / u1 - 1).max() <= 1e-12


def test_1p_pointsource_reference_vtu():
    """test_1p_pointsources_timeindependent_tpfa -> test_1p_pointsources_timeindependent_cc-reference.vtu: 10 kg/s at a grid vertex,
    shared by the four cells around it (pointsource.hh); pins the source term of the local residual (fvlocalresidual.hh:319-333)"""
    spec = problems.onep_pointsource()
    assert int((spec.source[:, 0] > 0).sum()) == 4
    u, st, rep = Oracle(spec).newton(spec.initial, spec.initial)
    assert st == 0
    g = np.load(os.path.join(GOLDEN, "test_1p_pointsources_timeindependent_cc.npz"))["p"].astype(np.float64)
    assert _fuzzy_ok(u, g)
    assert np.abs(u / g - 1).max() < 1e-5          # Float32 storage of the file
    assert u.max() > 1.6e5 and np.argmax(u) in np.flatnonzero(spec.source[:, 0] > 0)


def test_1p_incompressible_analytic_reference_vtu():
    """test_1p_incompressible_tpfa (DiffMethod::analytic, the reference's default for this test) -> the same
    test_1p_cc-reference.vtu; the analytic Jacobian (1p/incompressiblelocalresidual.hh:76-123,204-221) is exact, so it equals the
    FD Jacobian of the linear problem up to the FD rounding."""
    spec = problems.onep_incompressible((10, 10), analytic=True)
    x, o, its = _one_linear_step(spec)
    g = np.load(os.path.join(GOLDEN, "test_1p_cc.npz"))["p"].astype(np.float64)
    assert _fuzzy_ok(x, g) and np.abs(x / g - 1).max() < 5e-6
    _, ja = o.assemble(np.zeros(100))
    _, jn = Oracle(problems.onep_incompressible((10, 10), numdiff_params=True)).assemble(np.zeros(100))
    assert np.abs(ja - jn).max() <= 1e-9 * np.abs(ja).max()
    # closed form on the 10x10 grid: interior face tij = K (h = 0.1, area 0.1, two half distances 0.05), up = rho/mu = 1e6
    K = 1e-10
    assert ja[o.rowptr[0] + 1] == pytest.approx(-K * 1e6, rel=1e-14)           # A[0][1]
    # cell 0 touches the Dirichlet bottom (t = area*K/0.05 = 2K) and two interior faces: diagonal = (2K + K + K) * up
    assert ja[o.rowptr[0]] == pytest.approx(4 * K * 1e6, rel=1e-14)
    # row sums vanish away from Dirichlet faces (pure flux balance, no storage term)
    I = 45
    assert abs(ja[o.rowptr[I]:o.rowptr[I + 1]].sum()) <= 1e-12 * abs(ja[o.rowptr[I]:o.rowptr[I + 1]]).max()


def test_1p_analytic_and_numeric_jacobian_agree():
    """For the linear 1p problem the FD Jacobian is exact for any eps (test_1p_incompressible_tpfa vs _numdiff)."""
    a = problems.onep_incompressible((10, 10), numdiff_params=True)
    b = problems.onep_incompressible((10, 10), numdiff_params=False)
    xa, oa, _ = _one_linear_step(a)
    _, ja = oa.assemble(np.zeros(100))
    # with the default step eps = 1e-10*(|x|+1) the quotient at x = 0 loses most digits to cancellation (which is why the
    # reference's numdiff test overrides BaseEpsilon/PriVarMagnitude): the Jacobians agree only roughly
    _, jb = Oracle(b).assemble(np.zeros(100))
    assert np.abs(ja - jb).max() > 1e-9 * np.abs(ja).max()
    assert np.abs(ja - jb).max() < 0.5 * np.abs(ja).max()
    # closed form: d/dp of tij*(pI-pJ)*rho/mu = tij*rho/mu; interior homogeneous face of the 10x10 grid: tij = K
    K = 1e-10
    off = ja[oa.rowptr[0] + 1]                   # cell 0: columns (0, 1, 10) -> entry (0,1)
    assert off == pytest.approx(-K * 1000.0 / 1e-3, rel=1e-9)


def test_lognormal_permeability_matches_reference_field():
    """std::mt19937(0) + Dumux::SimpleLogNormalDistribution replay (examples/1ptracer/spatialparams_1p.hh:95-104)
    against the `permeability` field stored in test_1ptracer_pressure-reference.vtu."""
    spec = problems.onep_tracer_pressure((50, 50))
    g = np.load(os.path.join(GOLDEN, "test_1ptracer_pressure.npz"))
    ref = g["permeability"].astype(np.float64)
    assert np.abs(spec.K / ref - 1).max() < 5e-6          # 6 significant digits in the VTU


def test_1ptracer_pressure_50x50_reference_vtu():
    """examples/1ptracer stationary 1p solve on the heterogeneous field -> test_1ptracer_pressure-reference.vtu"""
    x, o, its = _one_linear_step(problems.onep_tracer_pressure((50, 50)))
    g = np.load(os.path.join(GOLDEN, "test_1ptracer_pressure.npz"))["p"].astype(np.float64)
    assert _fuzzy_ok(x, g)
    assert np.abs(x / g - 1).max() < 5e-6


@pytest.fixture(scope="module")
def lens_run():
    spec = problems.twop_lens((48, 32), law="vg")
    o = Oracle(spec)
    u, nsteps, its, dts = o.run_timeloop(spec.initial, 3000.0, 250.0)
    return spec, o, u.reshape(-1, 2), nsteps, its, dts


def test_2p_lens_48x32_reference_vtu(lens_run):
    """test_2p_incompressible_tpfa (van Genuchten lens, tEnd 3000 s, dt0 250 s) -> test_2p_incompressible_cc-reference.vtu.
    All ten stored fields are compared at the reference's fuzzy tolerance."""
    spec, o, u, nsteps, its, dts = lens_run
    # the golden file is output number 7 of the reference run (CMakeLists.txt: test_2p_incompressible_tpfa-00007.vtu)
    assert nsteps == 7 and abs(sum(dts) - 3000.0) < 1e-6
    g = np.load(os.path.join(GOLDEN, "test_2p_incompressible_cc.npz"))
    vv = o.volvars(u.reshape(-1))
    cols = {"S_aq": 0, "S_napl": 1, "p_aq": 2, "p_napl": 3, "rho_aq": 4, "rho_napl": 5, "mob_aq": 6, "mob_napl": 7, "pc": 8,
            "porosity": 9}
    for name, c in cols.items():
        ref = g[name].astype(np.float64)
        assert _fuzzy_ok(vv[:, c], ref), name
    # far tighter than the reference's own bar (Float32 storage is the limit): saturations to 5e-6 absolute, pressures to 1e-5 relative
    assert np.abs(u[:, 1] - g["S_napl"]).max() < 5e-6
    assert np.abs(u[:, 0] / g["p_aq"] - 1).max() < 1e-5


def test_2p_lens_analytic_jacobian_reference_vtu(lens_run):
    """test_2p_incompressible_tpfa_analytic (DiffMethod::analytic, 2p/incompressiblelocalresidual.hh:80-234,420-481; ILU0-GMRes) is
    compared by the reference against the SAME golden file, again output number 7: seven time steps, the Newton counts of the
    numeric run, S_n to Float32 precision.  The analytic Jacobian equals a central-difference one except where the finite
    difference straddles an upwind switch."""
    import dataclasses
    spec = problems.twop_lens((48, 32), law="vg", analytic=True)
    o = Oracle(spec)
    o.set_linear_solver("gmres", 10)
    u, nsteps, its, dts = o.run_timeloop(spec.initial, 3000.0, 250.0)
    assert nsteps == 7 and list(its) == list(lens_run[4])
    g = np.load(os.path.join(GOLDEN, "test_2p_incompressible_cc.npz"))
    u2 = u.reshape(-1, 2)
    assert np.abs(u2[:, 1] - g["S_napl"]).max() < 5e-6
    assert np.abs(u2[:, 0] / g["p_aq"] - 1).max() < 1e-5
    rng = np.random.RandomState(0)
    small = problems.twop_lens((24, 16), law="vg", analytic=True)
    cur = small.initial.copy()
    cur[:, 0] += rng.uniform(-50, 50, cur.shape[0])
    cur[:, 1] = rng.uniform(0.02, 0.6, cur.shape[0])
    prev = small.initial.copy()
    prev[:, 1] = rng.uniform(0, 0.3, cur.shape[0])
    ra, ja = Oracle(small).assemble(cur, prev)
    central = dataclasses.replace(small, options=dataclasses.replace(small.options, fd_method=0, base_eps=1e-6))
    rc, jc = Oracle(central).assemble(cur, prev)
    assert np.array_equal(ra, rc)
    rel = np.abs(ja - jc).reshape(-1, 4) / np.abs(jc).reshape(-1, 4).max(axis=0)
    assert np.mean(rel.max(axis=1) < 1e-7) > 0.99 and np.median(rel) < 1e-10


def test_2p_oilwet_lens_reference_vtu():
    """test_2p_incompressible_tpfa_oilwet (SpatialParams.LensIsOilWet, no gravity, DtInitial 130, ILURestartedGMResIstlSolver as in
    main.cc:134) -> test_2p_incompressible_tpfa_oilwet-reference.vtu, which is the NINTH output file: the run must reach
    t = 3000 s in exactly nine time steps, which pins the per-region wetting phase of TwoPVolumeVariables
    (2p/volumevariables.hh:87-96,132-152), the Newton counts and the time-step control together.  (ILU0-BiCGSTAB breaks down in
    the very first Newton iteration of this run -- rho == 0 exactly, the non-wetting rows are solved by the ILU sweep -- which
    is why the reference's test uses GMRes.)"""
    spec = problems.twop_lens((48, 32), law="vg", oilwet=True, dt=130.0)
    o = Oracle(spec)
    o.set_linear_solver("gmres", 10)
    u, nsteps, its, dts = o.run_timeloop(spec.initial, 3000.0, 130.0)
    assert nsteps == 9 and dts[0] == 130.0 and abs(sum(dts) - 3000.0) < 1e-9
    g = np.load(os.path.join(GOLDEN, "test_2p_incompressible_tpfa_oilwet.npz"))
    vv = o.volvars(u)
    cols = {"S_aq": 0, "S_napl": 1, "p_aq": 2, "p_napl": 3, "rho_aq": 4, "rho_napl": 5, "mob_aq": 6, "mob_napl": 7, "pc": 8,
            "porosity": 9}
    for name, c in cols.items():
        assert _fuzzy_ok(vv[:, c], g[name].astype(np.float64)), name
    # far tighter than the reference's bar: Float32 storage is the limit
    u2 = u.reshape(-1, 2)
    assert np.abs(u2[:, 1] - g["S_napl"]).max() < 5e-6
    assert np.abs(u2[:, 0] / g["p_aq"] - 1).max() < 2e-5
    assert np.abs(vv[:, 8] - g["pc"]).max() < 0.05                 # Pa, of up to 2530
    # inside the oil-wet lens the non-wetting (water) pressure exceeds the oil pressure: p_napl = p_aq - pc
    lens = spec.region == 1
    assert np.all(vv[lens, 3] <= vv[lens, 2]) and np.all(vv[~lens, 3] >= vv[~lens, 2])
    # BiCGSTAB on the same run: breakdown in the first linear solve -> the time step is halved until it is tiny
    ob = Oracle(spec)
    ub, nb, itsb, dtsb = ob.run_timeloop(spec.initial, 3000.0, 130.0)
    assert dtsb[0] < 130.0


def test_2p_lens_newton_counts_are_stable(lens_run):
    """The time-step / Newton control (newtonsolver.hh:784-798, timeloop.hh) is deterministic: pin the sequence so that a
    change in the oracle's control flow is noticed."""
    spec, o, u, nsteps, its, dts = lens_run
    assert dts[0] == 250.0
    assert all(2 <= n <= 18 for n in its)
    assert all(d2 >= d1 for d1, d2 in zip(dts[:-2], dts[1:-1]))      # dt only grows in this easy run
    # mass balance of the non-wetting phase: injected mass = stored mass (no TCE leaves through the Dirichlet sides yet)
    cells = spec.cells
    vol = (6.0 / cells[0]) * (4.0 / cells[1])
    stored = (0.4 * 1460.0 * u[:, 1] * vol).sum()
    injected = 0.04 * 1.0 * 3000.0            # 0.04 kg/(m^2 s) over the 1 m wide inlet
    assert stored == pytest.approx(injected, rel=2e-6)


def test_std_pow_switch_quantifies_fidelity():
    """SURVEY hard part 1: deviation of the Jacobian when glibc pow replaces the shared deterministic pow."""
    spec = problems.twop_lens((24, 16), law="bc")
    rng = np.random.RandomState(0)
    cur = spec.initial.copy()
    cur[:, 1] = rng.uniform(0.0, 0.3, size=cur.shape[0])
    r1, j1 = Oracle(spec).assemble(cur, spec.initial)
    r2, j2 = Oracle(spec, use_std_pow=True).assemble(cur, spec.initial)
    assert np.abs(r1 - r2).max() <= 1e-12 * np.abs(r1).max()
    # FD amplification (1/eps = 1e10) of last-ulp pow differences stays below 1e-4 relative of the row scale
    assert np.abs(j1 - j2).max() <= 1e-4 * np.abs(j1).max()
