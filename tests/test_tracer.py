"""Tracer transport on a frozen velocity field (BASELINE config 5, examples/1ptracer).

CPU: the oracle's volume fluxes + explicit tracer assembly reproduce the reference's own regression file
test/references/test_1ptracer_transport-reference.vtu (500 explicit Euler steps of 10 s on the 50x50 log-normal field).
GPU: volume fluxes, residual and Jacobian bit-identical to the oracle (explicit and implicit), and the same time loop
through the C ABI."""
import os

import numpy as np
import pytest

from dumux_b200 import problems
from oracle.oracle_py import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pressure_and_flux(cells):
    ps = problems.onep_tracer_pressure(cells)
    o = Oracle(ps)
    n = int(np.prod(cells))
    r, j = o.assemble(np.zeros(n))
    dx, st, its, red = o.solve(j, r, reduction=1e-13, maxit=1000)
    assert st == 0
    p = 0.0 - dx
    return ps, p, o.volume_flux(p)


def test_oracle_volume_flux_is_conservative_and_antisymmetric():
    ps, p, vf = _pressure_and_flux((20, 20))
    assert np.abs(vf.sum(axis=1)).max() <= 1e-7 * np.abs(vf).max()          # stationary incompressible: div q = 0 to solver tolerance
    v = vf.reshape(20, 20, 4)
    assert np.allclose(v[:, :-1, 1], -v[:, 1:, 0], rtol=1e-12, atol=0)     # +x of I  == -(-x of J)
    assert np.allclose(v[:-1, :, 3], -v[1:, :, 2], rtol=1e-6, atol=0)             # gravity terms cancel ~1e5 Pa of hydrostatic head
    assert np.all(v[:, 0, 0] == 0) and np.all(v[:, -1, 1] == 0)            # Neumann sides are skipped
    assert np.all(v[-1, :, 3] > 0) and np.all(v[0, :, 2] < 0)              # outflow at the top, inflow at the bottom


def test_oracle_tracer_transport_matches_reference_vtu():
    ps, p, vf = _pressure_and_flux((50, 50))
    ts = problems.tracer_transport((50, 50), vf, dt=10.0)
    ot = Oracle(ts)
    x = ts.initial.ravel().copy()
    for _ in range(500):                                                   # TEnd 5000 s, DtInitial = MaxTimeStepSize = 10 s
        r, j = ot.assemble(x, x)
        dx, st, its, red = ot.solve(j, r, reduction=1e-13)
        assert st == 0 and its <= 1                                        # diagonal system
        x = x - dx
    g = np.load(os.path.join(GOLDEN, "test_1ptracer_transport.npz"))
    X = g["X_tracer_0"].astype(np.float64)
    assert np.linalg.norm(x - X) <= 1e-5 * np.linalg.norm(X)               # Float32 storage; reference bar is 1e-2
    assert np.abs(x - X).max() <= 1e-5 * X.max()
    assert np.all(g["rho"] == 1000.0)


def test_oracle_implicit_tracer_is_consistent_with_its_jacobian():
    ps, p, vf = _pressure_and_flux((12, 10))
    ts = problems.tracer_transport((12, 10), vf, dt=50.0, implicit=True)
    ot = Oracle(ts)
    rng = np.random.RandomState(1)
    x = rng.uniform(0, 1e-9, size=120)
    prev = rng.uniform(0, 1e-9, size=120)
    r0, j = ot.assemble(x, prev)
    # the residual is linear in X: r(x + d) - r(x) == J d exactly up to rounding
    d = rng.uniform(-1e-10, 1e-10, size=120)
    r1, _ = ot.assemble(x + d, prev)
    Jd = np.zeros(120)
    for i in range(120):
        for k in range(ot.rowptr[i], ot.rowptr[i + 1]):
            Jd[i] += j[k] * d[ot.colidx[k]]
    assert np.abs((r1 - r0) - Jd).max() <= 1e-12 * np.abs(Jd).max()


CONSTVEL = [(False, "test_tracer_explicit_tpfa", 1e-8, "X_tracer_0"), (False, "test_tracer_explicit_tpfa", 0.0, "X_tracer_1"),
            (True, "test_tracer_implicit_tpfa", 1e-8, "X_tracer_0"), (True, "test_tracer_implicit_tpfa", 0.0, "X_tracer_1")]


@pytest.mark.parametrize("implicit,golden,D,field", CONSTVEL)
def test_oracle_constvel_tracer_matches_reference_vtu(implicit, golden, D, field):
    """test/porousmediumflow/tracer/constvel (test_tracer_explicit_tpfa / test_tracer_implicit_tpfa): 100 steps of 1e4 s on the
    analytic divergence-free velocity field.  The files hold two decoupled components: Problem.D = 1e-8 (Fick's law with
    DiffusivityConstantTortuosity, flux/cctpfa/fickslaw.hh) and D2 = 0 (pure advection); both must be reproduced by the explicit
    AND the implicit TracerLocalResidual restatement (the implicit assembler and Fick's law have no other golden).  Mass is
    conserved (all boundaries no-flow)."""
    ts = problems.tracer_constvel((50, 50), implicit=implicit, D=D)
    assert np.abs(ts.volume_flux.sum(axis=1)).max() <= 1e-4 * np.abs(ts.volume_flux).max()      # discretely divergence-free
    o = Oracle(ts)
    x = ts.initial.ravel().copy()
    for _ in range(100):
        r, j = o.assemble(x, x)
        dx, st, its, red = o.solve(j, r, reduction=1e-13, maxit=500)
        assert st == 0
        x = x - dx
    g = np.load(os.path.join(GOLDEN, golden + ".npz"))
    X = g[field].astype(np.float64)
    assert np.abs(x - X).max() <= 1e-5 * X.max() and np.linalg.norm(x - X) <= 5e-6 * np.linalg.norm(X)
    assert x.sum() == pytest.approx(ts.initial.sum(), rel=1e-11)
    # diffusion matters: the two components of the file differ visibly
    assert np.abs(g["X_tracer_0"] - g["X_tracer_1"]).max() > 0.1 * g["X_tracer_1"].max()


DISPERSION = [(1e-8, "X_tracer_0"), (0.0, "X_tracer_1")]


def _constvel_run(stepper_assemble_solve, ts, steps=100):
    x = ts.initial.ravel().copy()
    for _ in range(steps):
        x = stepper_assemble_solve(x)
    return x


@pytest.mark.parametrize("D,field", DISPERSION)
def test_oracle_constvel_dispersion_matches_reference_vtu(D, field):
    """test_tracer_implicit_dispersion_tpfa (-Problem.AlphaL 0.02 -Problem.AlphaT 0.008): mechanical dispersion with Scheidegger's
    tensor from the analytic velocity field (dispersiontensors/scheidegger.hh, flux/cctpfa/dispersionflux.hh), implicit assembler
    with the reference's analytic Jacobian, which has no dispersion derivative (one linear solve per time step, main.cc) -> 100
    steps, both components of test_tracer_implicit_dispersion_tpfa-reference.vtu to the file's Float32 precision; without the
    dispersion term the fields differ by 20 % / 43 %"""
    ts = problems.tracer_constvel((50, 50), implicit=True, D=D, alpha_l=0.02, alpha_t=0.008)
    assert ts.tracer_dispersion.min() >= 0.0 and ts.tracer_dispersion.max() > 0.0
    o = Oracle(ts)

    def step(x):
        r, j = o.assemble(x, x)
        dx, st, its, red = o.solve(j, r, reduction=1e-13, maxit=500)
        assert st == 0
        return x - dx

    x = _constvel_run(step, ts)
    X = np.load(os.path.join(GOLDEN, "test_tracer_implicit_dispersion_tpfa.npz"))[field].astype(np.float64)
    assert np.abs(x - X).max() <= 1e-5 * X.max() and np.linalg.norm(x - X) <= 5e-6 * np.linalg.norm(X)
    assert x.sum() == pytest.approx(ts.initial.sum(), rel=1e-11)
    plain = np.load(os.path.join(GOLDEN, "test_tracer_implicit_tpfa.npz"))[field].astype(np.float64)
    assert np.linalg.norm(plain - X) > 0.1 * np.linalg.norm(X)


def test_oracle_dispersion_enters_the_residual_only():
    """the Jacobian of the implicit assembler is the reference's analytic one (advection + Fick), unchanged by the dispersion
    array; the residual differs by the dispersive flux, which is conservative"""
    a = problems.tracer_constvel((12, 9), implicit=True, D=3e-7)
    b = problems.tracer_constvel((12, 9), implicit=True, D=3e-7, alpha_l=0.02, alpha_t=0.008)
    rng = np.random.RandomState(4)
    x, prev = rng.uniform(0, 1e-9, size=108), rng.uniform(0, 1e-9, size=108)
    ra, ja = Oracle(a).assemble(x, prev)
    rb, jb = Oracle(b).assemble(x, prev)
    assert np.array_equal(ja, jb) and not np.array_equal(ra, rb)
    assert abs((rb - ra).sum()) <= 1e-12 * np.abs(rb - ra).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("D,field", DISPERSION)
def test_gpu_constvel_dispersion_matches_oracle_and_reference_vtu(engine_factory, D, field):
    from dumux_b200 import binding as B
    ts = problems.tracer_constvel((50, 50), implicit=True, D=D, alpha_l=0.02, alpha_t=0.008)
    o = Oracle(ts)
    rng = np.random.RandomState(2)
    x0, p0 = rng.uniform(0, 1e-9, size=2500), rng.uniform(0, 1e-9, size=2500)
    e = engine_factory(ts)
    rg, jg = e.assemble(x0, p0)
    ro, jo = o.assemble(x0, p0)
    assert np.array_equal(rg, ro) and np.array_equal(jg, jo)          # dispersive flux bit-identical
    e.upload(B.VEC_CUR, ts.initial)
    e.upload(B.VEC_PREV, ts.initial)
    prm = e.newton_params(lin_reduction=1e-13, lin_maxit=500)
    for _ in range(100):
        st, its, shift, a, s_, u = e.newton_step(prm)
        assert st == 0
        e.advance_timestep()
    x = e.download(B.VEC_CUR).ravel()
    X = np.load(os.path.join(GOLDEN, "test_tracer_implicit_dispersion_tpfa.npz"))[field].astype(np.float64)
    assert np.abs(x - X).max() <= 1e-5 * X.max() and np.linalg.norm(x - X) <= 5e-6 * np.linalg.norm(X)


def test_oracle_fick_jacobian_is_the_derivative_of_the_implicit_residual():
    ts = problems.tracer_constvel((12, 9), implicit=True, D=3e-7)
    o = Oracle(ts)
    rng = np.random.RandomState(2)
    cur, prev = rng.uniform(0, 2e-11, size=(108, 1)), rng.uniform(0, 2e-11, size=(108, 1))
    r0, j = o.assemble(cur, prev)
    import scipy.sparse as sp
    A = sp.csr_matrix((j, o.colidx, o.rowptr), shape=(108, 108))
    d = rng.uniform(-1e-12, 1e-12, size=(108, 1))
    r1, _ = o.assemble(cur + d, prev, jacobian=False)
    assert np.linalg.norm((r1 - r0) - A @ d.ravel()) <= 1e-9 * np.linalg.norm(A @ d.ravel())      # the residual is linear in X


@pytest.mark.gpu
@pytest.mark.parametrize("implicit,golden,D,field", CONSTVEL)
def test_gpu_constvel_tracer_matches_reference_vtu(engine_factory, implicit, golden, D, field):
    """The same 100 steps on the device (state resident; explicit steps take the diagonal fast path of the ILU0 factorisation),
    and the assembly with Fick's law bit-identical to the oracle."""
    from dumux_b200 import binding as B
    ts = problems.tracer_constvel((50, 50), implicit=implicit, D=D)
    rng = np.random.RandomState(8)
    cur, prev = rng.uniform(0, 2e-11, size=(2500, 1)), rng.uniform(0, 2e-11, size=(2500, 1))
    ro, jo = Oracle(ts).assemble(cur, prev)
    et = engine_factory(ts)
    rg, jg = et.assemble(cur, prev)
    assert np.array_equal(rg, ro) and np.array_equal(jg, jo)
    et.upload(B.VEC_CUR, ts.initial)
    et.upload(B.VEC_PREV, ts.initial)
    prm = et.newton_params(lin_reduction=1e-13, lin_maxit=500)
    for _ in range(100):
        st, its, shift, a, s, upd = et.newton_step(prm)
        assert st == 0
        et.advance_timestep()
    x = et.download(B.VEC_CUR).ravel()
    X = np.load(os.path.join(GOLDEN, golden + ".npz"))[field].astype(np.float64)
    assert np.abs(x - X).max() <= 1e-5 * X.max() and np.linalg.norm(x - X) <= 5e-6 * np.linalg.norm(X)


@pytest.mark.gpu
@pytest.mark.parametrize("cells", [(50, 50), (21, 13, 17)])
def test_gpu_volume_flux_bit_identical(engine_factory, cells):
    ps = problems.onep_tracer_pressure(cells)
    rng = np.random.RandomState(3)
    p = 1e5 + rng.uniform(0, 1e4, size=int(np.prod(cells)))
    vo = Oracle(ps).volume_flux(p)
    vg = engine_factory(ps).volume_flux(p)
    assert np.array_equal(vg, vo)


@pytest.mark.gpu
@pytest.mark.parametrize("implicit", [False, True])
@pytest.mark.parametrize("cells", [(50, 50), (21, 13, 17)])
def test_gpu_tracer_assembly_bit_identical(engine_factory, cells, implicit):
    ps = problems.onep_tracer_pressure(cells)
    rng = np.random.RandomState(4)
    n = int(np.prod(cells))
    vf = Oracle(ps).volume_flux(1e5 + rng.uniform(0, 1e4, size=n))
    ts = problems.tracer_transport(cells, vf, dt=10.0, implicit=implicit)
    cur = rng.uniform(0, 2e-11, size=(n, 1))
    prev = rng.uniform(0, 2e-11, size=(n, 1))
    ro, jo = Oracle(ts).assemble(cur, prev)
    e = engine_factory(ts)
    rg, jg = e.assemble(cur, prev)
    assert np.array_equal(rg, ro) and np.array_equal(jg, jo)
    r2, j2 = e.assemble(cur, prev, jacobian=False)
    assert j2 is None and np.array_equal(r2, ro)


@pytest.mark.gpu
def test_gpu_1ptracer_example_end_to_end(engine_factory):
    """examples/1ptracer/main.cc: stationary 1p solve -> volume fluxes -> 500 explicit tracer steps, all on the device."""
    from dumux_b200 import binding as B
    ps = problems.onep_tracer_pressure((50, 50))
    e1 = engine_factory(ps)
    u, st, rep = e1.newton(np.zeros((2500, 1)), np.zeros((2500, 1)), min_steps=1, max_steps=1, lin_reduction=1e-13, lin_maxit=1000,
                           max_rel_shift=1e300)
    vf = e1.volume_flux(u)
    ts = problems.tracer_transport((50, 50), vf, dt=10.0)
    et = engine_factory(ts)
    et.upload(B.VEC_CUR, ts.initial)
    et.upload(B.VEC_PREV, ts.initial)
    prm = et.newton_params(lin_reduction=1e-13)
    for _ in range(500):
        st, its, shift, a, s, upd = et.newton_step(prm)
        assert st == 0
        et.advance_timestep()
    x = et.download(B.VEC_CUR).ravel()
    X = np.load(os.path.join(GOLDEN, "test_1ptracer_transport.npz"))["X_tracer_0"].astype(np.float64)
    assert np.linalg.norm(x - X) <= 1e-5 * np.linalg.norm(X)
