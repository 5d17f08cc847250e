"""Diagonal permeability tensors (SpatialParams::permeability returning a FieldMatrix without off-diagonal entries;
dmx_set_permeability_diagonal / orc_set_permeability_diagonal) and the reference test that needs them:
test/porousmediumflow/1p/convergence/analyticsolution, variant test_1p_convergence_analytic_tpfa_structured (Problem.C = 0 ->
K = diag(1, exp(-2))): refinements 0..3 of a 10 x 10 grid, accepted if the mean convergence rate of the discrete L2 pressure
error is >= 1.8 (convergencetest.py)."""
import dataclasses

import numpy as np
import pytest

from dumux_b200 import problems
from oracle.oracle_py import Oracle


def _rates(errs, ndofs):
    return [(np.log(errs[i]) - np.log(errs[i + 1])) / np.log(np.sqrt(ndofs[i + 1] / ndofs[i])) for i in range(len(errs) - 1)]


def _aniso(spec, factors, seed=1):
    """the spec with K -> K * factors[a] * (1 + noise_a) per axis"""
    rng = np.random.RandomState(seed)
    K = np.asarray(spec.K, dtype=np.float64)
    Kd = np.stack([K * f * rng.uniform(0.5, 1.5, size=K.shape[0]) for f in factors[:spec.dim]], axis=1)
    return dataclasses.replace(spec, K=Kd)


def test_convergence_analytic_tpfa_structured_oracle():
    errs, nd = [], []
    for ref in range(4):
        N = 10 * 2 ** ref
        spec = problems.onep_convergence((N, N))
        u, st, rep = Oracle(spec).newton(spec.initial, spec.initial)
        assert st == 0
        errs.append(problems.onep_convergence_l2_error(spec, u))
        nd.append(N * N)
    rates = _rates(errs, nd)
    assert np.mean(rates) >= 1.8 and all(r > 1.9 for r in rates), rates
    # the tensor matters: with the scalar K = K_xx the discrete solution does not converge to the analytic one
    spec = problems.onep_convergence((40, 40))
    iso = dataclasses.replace(spec, K=np.ascontiguousarray(spec.K[:, 0]))
    u, st, rep = Oracle(iso).newton(iso.initial, iso.initial)
    assert problems.onep_convergence_l2_error(iso, u) > 100 * errs[2]


@pytest.mark.parametrize("make", [lambda: problems.twop_lens((9, 7, 6), law="bc", heterogeneity_sigma=0.3),
                                  lambda: problems.onep_compressible((8, 7, 5), lognormal=True),
                                  lambda: problems.twop_lens((12, 9), law="vg")])
def test_equal_diagonal_entries_are_the_scalar_field(make):
    """K = diag(k, k, k) is the scalar permeability k: bit-identical residual and Jacobian"""
    spec = make()
    o = Oracle(spec)
    rng = np.random.RandomState(2)
    cur = spec.initial.copy()
    cur[:, 0] += rng.uniform(-50, 50, size=cur.shape[0])
    r0, j0 = o.assemble(cur, spec.initial)
    K = np.asarray(spec.K)
    o2 = Oracle(dataclasses.replace(spec, K=np.stack([K] * spec.dim, axis=1)))
    r1, j1 = o2.assemble(cur, spec.initial)
    assert np.array_equal(r0, r1) and np.array_equal(j0, j1)
    # ... and a genuinely anisotropic field changes the fluxes along the weaker axes
    o3 = Oracle(_aniso(spec, (1.0, 0.3, 0.05)))
    r3, j3 = o3.assemble(cur, spec.initial)
    assert not np.array_equal(j0, j3)


@pytest.mark.gpu
def test_convergence_analytic_tpfa_structured_device(engine_factory):
    errs, nd = [], []
    for ref in range(4):
        N = 10 * 2 ** ref
        spec = problems.onep_convergence((N, N))
        uo, sto, repo = Oracle(spec).newton(spec.initial, spec.initial)
        e = engine_factory(spec)
        ug, stg, repg = e.newton(spec.initial, spec.initial)
        assert sto == 0 and stg == 0 and repg.newton_iterations == repo.newton_iterations
        assert np.linalg.norm(ug - uo) <= 1e-8 * np.linalg.norm(uo)
        errs.append(problems.onep_convergence_l2_error(spec, ug))
        nd.append(N * N)
        e.close()
    rates = _rates(errs, nd)
    assert np.mean(rates) >= 1.8, rates


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2p-3d", "2p-3d-analytic", "1p-table-3d", "2p-2d", "1p-2d-analytic"])
def test_assembly_bit_exact_with_diagonal_tensor(engine_factory, name):
    """transmissibilities, Dirichlet faces, gravity terms (entry along the gravity axis) and the analytic Jacobians with K_aa per axis"""
    spec = {"2p-3d": lambda: problems.twop_lens((13, 9, 7), law="bc", heterogeneity_sigma=0.4),
            "2p-3d-analytic": lambda: problems.twop_lens((13, 9, 7), law="bc", heterogeneity_sigma=0.4, analytic=True),
            "1p-table-3d": lambda: problems.onep_compressible((9, 8, 7), lognormal=True),
            "2p-2d": lambda: problems.twop_lens((24, 16), law="vg"),
            "1p-2d-analytic": lambda: problems.onep_incompressible((12, 10), analytic=True)}[name]()
    spec = _aniso(spec, (1.0, 0.3, 0.05))
    o = Oracle(spec)
    rng = np.random.RandomState(5)
    cur = spec.initial.copy()
    cur[:, 0] += rng.uniform(-80, 80, size=cur.shape[0])
    if spec.num_eq == 2:
        cur[:, 1] = rng.uniform(0.0, 0.3, size=cur.shape[0])
    ro, jo = o.assemble(cur, spec.initial)
    e = engine_factory(spec)
    rg, jg = e.assemble(cur, spec.initial)
    assert np.array_equal(ro, rg)
    assert np.array_equal(jo, jg)
    e.close()


@pytest.mark.gpu
def test_volume_flux_with_diagonal_tensor(engine_factory):
    ps = _aniso(problems.onep_tracer_pressure((14, 11, 9)), (1.0, 0.4, 0.1))
    rng = np.random.RandomState(3)
    ctr = problems.cell_centers(ps.cells, ps.lower, ps.upper)
    p = 1.0e5 * (1.1 - 0.1 * ctr[:, 2]) + rng.uniform(-20.0, 20.0, size=ctr.shape[0])
    vo = Oracle(ps).volume_flux(p)
    e = engine_factory(ps)
    vg = e.volume_flux(p)
    assert np.array_equal(vo, vg)
    e.close()
