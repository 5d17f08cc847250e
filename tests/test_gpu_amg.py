"""AMG preconditioner (AMGBiCGSTABIstlSolver / AMGCGIstlSolver, dumux/linear/istlsolvers.hh:716-757) on the device against its
restatement oracle/amg_oracle.py: hierarchy, Galerkin coarse matrices, one V-cycle, preconditioned Krylov solves and a Newton
solve.  Every summation order is shared, so matrices and cycle outputs are compared bit for bit and iteration counts with ==.
dune-istl's own aggregation heuristic is not available here; the cycle and its defaults are dune's."""
import numpy as np
import pytest

from dumux_b200 import binding as B
from dumux_b200 import problems
from oracle import dist_oracle as D
from oracle import oracle_py as O
from oracle.amg_oracle import AmgOracle

pytestmark = pytest.mark.gpu

SPECS = {
    "2p-3d-odd": lambda: problems.twop_lens((13, 9, 7), law="bc", heterogeneity_sigma=0.4),
    "2p-3d": lambda: problems.twop_lens((20, 18, 22), law="bc", heterogeneity_sigma=0.5, plane_rng=True),
    "2p-2d": lambda: problems.twop_lens((24, 16), law="vg"),
    "1p-2d": lambda: problems.onep_incompressible((30, 20)),
}


def _system(spec, seed=0):
    o = O.Oracle(spec)
    rng = np.random.RandomState(seed)
    cur = spec.initial.copy()
    cur[:, 0] += rng.uniform(-100, 100, size=cur.shape[0])
    if spec.num_eq == 2:
        cur[:, 1] = rng.uniform(0, 0.3, size=cur.shape[0])
    res, jac = o.assemble(cur, spec.initial)
    return o, res, jac


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("smoother,pre,post", [("ssor", 2, 2), ("ilu", 1, 1), ("ssor", 0, 1)])
def test_hierarchy_galerkin_and_vcycle_bit_exact(engine_factory, name, smoother, pre, post):
    spec = SPECS[name]()
    o, res, jac = _system(spec)
    amg = AmgOracle(spec.cells, spec.dim, o.b, o.rowptr, o.colidx, jac, pre_steps=pre, post_steps=post, smoother=smoother)
    e = engine_factory(spec)
    e.set_amg_params(pre_steps=pre, post_steps=post, smoother=B.PRECOND_SSOR if smoother == "ssor" else B.PRECOND_ILU0)
    e.upload_jacobian(jac)
    e.upload(B.VEC_WORK0, res)
    e.precond_apply(B.PRECOND_AMG, B.VEC_WORK0, B.VEC_WORK1)
    assert e.amg_levels() == [lv.cells for lv in amg.levels] and len(amg.levels) >= 3
    for l in range(1, len(amg.levels)):
        assert np.array_equal(e.amg_level_matrix(l), amg.levels[l].values), l
    vg = e.download(B.VEC_WORK1)
    vo = amg.apply(res)
    assert np.array_equal(vg, vo)
    assert np.array_equal(e.download(B.VEC_WORK0), res)          # the defect handed in is not modified
    e.set_amg_params()


@pytest.mark.parametrize("name", ["2p-3d", "2p-3d-odd", "2p-2d"])
def test_amg_bicgstab_counts_equal_oracle(engine_factory, name):
    """AMGBiCGSTABIstlSolver: same iteration count as the oracle (device summation tree for the scalar products), far fewer
    iterations than ILU0-BiCGSTAB, same solution"""
    spec = SPECS[name]()
    ro = D.single_rank(spec)
    u0 = spec.initial.reshape(-1).copy()
    res, jac = ro.o.assemble(u0, u0)
    xo, sto, ito, redo = ro.bicgstab(jac, res, 1e-8, 500, precond="amg")
    xi, sti, iti, redi = ro.bicgstab(jac, res, 1e-8, 500, precond="ilu0")
    e = engine_factory(spec)
    xg, stg, itg, redg = e.solve(jac, res, reduction=1e-8, maxit=500, precond=B.PRECOND_AMG)
    assert sto == 0 and stg == 0 and itg == ito, (itg, ito)
    assert itg < iti
    assert np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)
    assert np.linalg.norm(xg - xi) <= 1e-6 * np.linalg.norm(xi)


def test_amg_cg_on_the_symmetric_1p_problem(engine_factory):
    """AMGCGIstlSolver (the solver of examples/1ptracer/main.cc:125-133 for the stationary pressure problem)"""
    spec = SPECS["1p-2d"]()
    ro = D.single_rank(spec)
    res, jac = ro.o.assemble(np.zeros(ro.n), None)
    xo, sto, ito, redo = ro.cg(jac, res, 1e-10, 500, precond="amg")
    e = engine_factory(spec)
    e.set_linear_solver("cg")
    xg, stg, itg, redg = e.solve(jac, res, reduction=1e-10, maxit=500, precond=B.PRECOND_AMG)
    e.set_linear_solver("bicgstab")
    assert sto == 0 and stg == 0 and itg == ito, (itg, ito)
    assert np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)


def test_newton_with_amg_matches_oracle(engine_factory):
    spec = problems.twop_lens((24, 20, 18), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
    ro = D.single_rank(spec)
    uo, sto, nsteps_o, lin_o = ro.newton(spec.initial, spec.initial, precond="amg")
    ui, sti, nsteps_i, lin_i = ro.newton(spec.initial, spec.initial, precond="ilu0")
    e = engine_factory(spec)
    ug, stg, rep = e.newton(spec.initial, spec.initial, preconditioner=B.PRECOND_AMG)
    lin_g = [rep.linear_iterations[i] for i in range(rep.newton_iterations)]
    assert sto == 0 and stg == 0 and rep.newton_iterations == nsteps_o == nsteps_i
    assert lin_g == lin_o, (lin_g, lin_o)
    assert sum(lin_g) * 4 < sum(lin_i)
    up, us = ug.reshape(-1, 2), uo.reshape(-1, 2)
    assert np.linalg.norm(up[:, 0] - us[:, 0]) <= 1e-8 * np.linalg.norm(us[:, 0])
    assert np.linalg.norm(up[:, 1] - us[:, 1]) <= 1e-8 * max(1.0, np.linalg.norm(us[:, 1]))
    # ... and the ILU0 solver still works on the same context afterwards (the smoother set-up shares its machinery)
    ug2, stg2, rep2 = e.newton(spec.initial, spec.initial)
    assert stg2 == 0 and [rep2.linear_iterations[i] for i in range(rep2.newton_iterations)] == lin_i


def test_amg_at_64_cubed_is_mesh_independent(engine_factory):
    spec = problems.twop_lens((64, 64, 64), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
    e = engine_factory(spec)
    e.upload(B.VEC_CUR, spec.initial)
    e.upload(B.VEC_PREV, spec.initial)
    st, its_amg, *_ = e.newton_step(e.newton_params(preconditioner=B.PRECOND_AMG))
    e.upload(B.VEC_CUR, spec.initial)
    st2, its_ilu, *_ = e.newton_step(e.newton_params(lin_maxit=2000))
    assert st == 0 and st2 == 0 and its_amg <= 8 and its_ilu > 60


# ------------------------------------------------------------------------------------------------------------------------
# test/linear/test_parallel_amg_smoothers.cc: Dune::Amg::AMG with Dumux::ParMTSSOR / ParMTSOR / ParMTJac as the smoother
# (SmootherArgs: 2 iterations, relaxation 0.8) preconditioning CG on the CCTpfa Helmholtz operator -div grad u + u of a 300 x 300
# YaspGrid (homogeneous Neumann), right-hand side b = A 1, reduction 1e-15, at most 200 iterations; accepted if the solver
# converged and | |x|^2 - N | <= 1e-10 N.  Same set-up here (the hierarchy is this library's).
# ------------------------------------------------------------------------------------------------------------------------
def helmholtz_system(N, o):
    """makeHelmholtzMatrix<CCTpfaModel>(gridView, a = 1, b = 1) on the unit square (linear/helmholtzoperator.hh:96-146): t_ij = 1 on a
    uniform 2-D grid, diagonal = number of neighbours + b h^2"""
    rp, ci = o.rowptr, o.colidx
    n, h = N * N, 1.0 / N
    rows = np.repeat(np.arange(n), np.diff(rp))
    vals = np.where(ci == rows, 0.0, -1.0)
    vals[ci == rows] = (np.diff(rp) - 1) + h * h
    return vals, O.spmv(n, 1, rp, ci, vals, np.ones(n))


PARMT = {"par_mt_ssor": B.PRECOND_PARMT_SSOR, "par_mt_sor": B.PRECOND_PARMT_SOR, "par_mt_jac": B.PRECOND_PARMT_JAC}


@pytest.mark.parametrize("smoother", list(PARMT))
def test_parallel_amg_smoothers_helmholtz(engine_factory, smoother):
    N = 300
    spec = problems.onep_incompressible((N, N))
    ro = D.single_rank(spec)
    vals, b = helmholtz_system(N, ro.o)
    e = engine_factory(spec)
    e.set_amg_params(smoother=PARMT[smoother], smoother_iterations=2, smoother_relaxation=0.8)
    e.set_linear_solver("cg")
    x, st, its, red = e.solve(vals, b, reduction=1e-15, maxit=200, precond=B.PRECOND_AMG)
    e.set_linear_solver("bicgstab")
    e.set_amg_params()
    assert st == 0, "Solver did not converge!"
    assert abs(np.dot(x, x) - x.size) <= 1e-10 * x.size
    # ... and it is the iteration the oracle runs (device summation tree for the scalar products)
    ro.amg_params = dict(smoother=smoother, smoother_iterations=2, smoother_relaxation=0.8)
    xo, sto, ito, redo = ro.cg(vals, b, 1e-15, 200, precond="amg")
    assert sto == 0 and abs(its - ito) <= 1, (its, ito)
    assert np.linalg.norm(x - xo) <= 1e-12 * np.linalg.norm(xo)


@pytest.mark.parametrize("smoother", list(PARMT))
def test_parmt_smoothed_vcycle_bit_exact(engine_factory, smoother):
    spec = SPECS["2p-3d-odd"]()
    o, res, jac = _system(spec)
    amg = AmgOracle(spec.cells, spec.dim, o.b, o.rowptr, o.colidx, jac, smoother=smoother, smoother_iterations=2, smoother_relaxation=0.8)
    e = engine_factory(spec)
    e.set_amg_params(smoother=PARMT[smoother], smoother_iterations=2, smoother_relaxation=0.8)
    e.upload_jacobian(jac)
    e.upload(B.VEC_WORK0, res)
    e.precond_apply(B.PRECOND_AMG, B.VEC_WORK0, B.VEC_WORK1)
    assert np.array_equal(e.download(B.VEC_WORK1), amg.apply(res))
    e.set_amg_params()
    with pytest.raises(B.DmxError):
        e.set_amg_params(smoother=B.PRECOND_SSOR, smoother_iterations=2)       # the factorised sweeps run one iteration
