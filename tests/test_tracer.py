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
