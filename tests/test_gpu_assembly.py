"""GPU parity: CUDA assembly (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): residual and Jacobian entries to a relative error of 1e-10 in fp64.
Relative error is measured against the largest magnitude in the same block row (entries that are zero in
both are exact); the residual is compared against the norm of the row's flux magnitudes via the max norm.
"""
import numpy as np
import pytest

from dumux_b200 import problems
from oracle.oracle_py import Oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _row_scale(rowptr, jac, n, b):
    """max |entry| per block row, broadcast to every entry of that row."""
    bb = b * b
    per_block = np.abs(jac.reshape(-1, bb)).max(axis=1)
    row_of_block = np.repeat(np.arange(n), np.diff(rowptr))
    row_max = np.zeros(n)
    np.maximum.at(row_max, row_of_block, per_block)
    return np.repeat(row_max[row_of_block], bb)


def _compare(spec, engine_factory, cur, prev):
    o = Oracle(spec)
    res_o, jac_o = o.assemble(cur, prev)
    e = engine_factory(spec)
    rp, ci = e.pattern()
    assert np.array_equal(rp, o.rowptr) and np.array_equal(ci, o.colidx)
    res_g, jac_g = e.assemble(cur, prev)
    scale = _row_scale(o.rowptr, jac_o, o.n, o.b)
    jerr = np.abs(jac_g - jac_o) / np.maximum(scale, 1e-300)
    rscale = max(np.abs(res_o).max(), 1e-300)
    rerr = np.abs(res_g - res_o).max() / rscale
    return rerr, jerr.max(), (res_o, jac_o, res_g, jac_g)


def _perturbed(spec, seed, dp=50.0, ds=0.3):
    rng = np.random.RandomState(seed)
    u = spec.initial.copy()
    u[:, 0] += rng.uniform(-dp, dp, size=u.shape[0])
    if spec.num_eq == 2:
        u[:, 1] = rng.uniform(0.0, ds, size=u.shape[0])
    return u


@pytest.mark.parametrize("cells", [(10, 10), (100, 100), (17, 9, 11)])
def test_1p_incompressible(engine_factory, cells):
    spec = problems.onep_incompressible(cells)
    cur = _perturbed(spec, 1, dp=1e4)
    rerr, jerr, _ = _compare(spec, engine_factory, cur, None)
    assert rerr <= 1e-13 and jerr <= RTOL, (rerr, jerr)


@pytest.mark.parametrize("cells", [(10, 10), (70, 45), (19, 12, 9)])
def test_1p_incompressible_analytic(engine_factory, cells):
    """DiffMethod::analytic (assembly/cclocalassembler.hh:490-600, 1p/incompressiblelocalresidual.hh:76-123,204-221): residual and
    Jacobian bit-identical to the oracle on a heterogeneous field."""
    spec = problems.onep_incompressible(cells, analytic=True)
    spec.K = spec.K * problems.fast_lognormal_multiplier(spec.num_cells, 0.5, 3)
    cur = _perturbed(spec, 1, dp=1e4)
    rerr, jerr, (res_o, jac_o, res_g, jac_g) = _compare(spec, engine_factory, cur, None)
    assert np.array_equal(res_g, res_o) and np.array_equal(jac_g, jac_o), (rerr, jerr)
    # one linear solve with the analytic Jacobian solves the (linear) problem: the residual at the new state vanishes
    e = engine_factory(spec)
    dx, st, its, red = e.solve(jac_g, res_g, reduction=1e-13, maxit=2000)
    assert st == 0
    res2, _ = e.assemble(cur - dx.reshape(cur.shape), None)
    assert np.linalg.norm(res2) <= 1e-9 * np.linalg.norm(res_g)


@pytest.mark.parametrize("law", ["vg", "bc"])
@pytest.mark.parametrize("cells", [(48, 32), (24, 12, 10)])
def test_2p_lens(engine_factory, law, cells):
    spec = problems.twop_lens(cells, law=law, heterogeneity_sigma=0.5 if len(cells) == 3 else 0.0)
    prev = _perturbed(spec, 2)
    cur = _perturbed(spec, 3)
    rerr, jerr, _ = _compare(spec, engine_factory, cur, prev)
    assert rerr <= 1e-13 and jerr <= RTOL, (rerr, jerr)


@pytest.mark.parametrize("law", ["vg", "bc"])
@pytest.mark.parametrize("cells", [(48, 32), (20, 12, 9)])
def test_2p_analytic_jacobian(engine_factory, law, cells):
    """DiffMethod::analytic for the incompressible 2p model (2p/incompressiblelocalresidual.hh:80-234,420-481): bit-identical to
    the oracle, saturations spanning the regularised branches."""
    spec = problems.twop_lens(cells, law=law, heterogeneity_sigma=0.5 if len(cells) == 3 else 0.0, analytic=True)
    rng = np.random.RandomState(12)
    cur = _perturbed(spec, 6)
    cur[:, 1] = rng.choice([-0.01, 0.0, 1e-9, 0.03, 0.3, 0.6, 0.9, 0.97, 1.0, 1.02], size=cur.shape[0])
    prev = _perturbed(spec, 7)
    rerr, jerr, (res_o, jac_o, res_g, jac_g) = _compare(spec, engine_factory, cur, prev)
    assert np.array_equal(res_g, res_o) and np.array_equal(jac_g, jac_o), (rerr, jerr)


@pytest.mark.parametrize("cells", [(48, 32), (20, 12, 9)])
def test_2p_oilwet_lens(engine_factory, cells):
    """Per-region wetting phase (2p/volumevariables.hh:87-96,132-152; test_2p_incompressible_tpfa_oilwet): in the lens phase 1
    wets, p1 = p0 - pc, krw belongs to phase 1.  Saturations span the regularised branches on both sides; bit-identical."""
    spec = problems.twop_lens(cells, law="vg", oilwet=True, dt=130.0)
    rng = np.random.RandomState(11)
    cur = _perturbed(spec, 4)
    cur[:, 1] = rng.choice([-0.01, 0.0, 1e-9, 0.03, 0.3, 0.6, 0.9, 0.97, 1.0, 1.02], size=cur.shape[0])
    prev = _perturbed(spec, 5, ds=0.9)
    rerr, jerr, (res_o, jac_o, res_g, jac_g) = _compare(spec, engine_factory, cur, prev)
    assert np.array_equal(res_g, res_o) and np.array_equal(jac_g, jac_o), (rerr, jerr)
    # the wetting phase matters: the same state with a water-wet lens gives another residual
    spec_ww = problems.twop_lens(cells, law="vg", oilwet=True, dt=130.0)
    spec_ww.materials[1] = spec_ww.materials[0]
    res_ww, _ = Oracle(spec_ww).assemble(cur, prev)
    assert np.abs(res_ww - res_o).max() > 1e-6 * np.abs(res_o).max()


def test_2p_saturation_extremes(engine_factory):
    """Regularised branches: S_n < 0, S_n = 0, S_w below the low-saturation threshold, S_w > 1."""
    spec = problems.twop_lens((16, 8, 6), law="vg")
    rng = np.random.RandomState(7)
    cur = spec.initial.copy()
    cur[:, 1] = rng.choice([-0.02, 0.0, 1e-12, 0.05, 0.5, 0.93, 0.995, 1.01], size=cur.shape[0])
    prev = spec.initial.copy()
    rerr, jerr, _ = _compare(spec, engine_factory, cur, prev)
    assert rerr <= 1e-13 and jerr <= RTOL, (rerr, jerr)
    spec = problems.twop_lens((16, 8, 6), law="bc")
    rerr, jerr, _ = _compare(spec, engine_factory, cur, prev)
    assert rerr <= 1e-13 and jerr <= RTOL, (rerr, jerr)


@pytest.mark.parametrize("method", [1, 0, -1, 5])
def test_fd_methods(engine_factory, method):
    spec = problems.twop_lens((12, 10, 8), law="bc", heterogeneity_sigma=0.3)
    spec.options.fd_method = method
    prev = _perturbed(spec, 4)
    cur = _perturbed(spec, 5)
    rerr, jerr, _ = _compare(spec, engine_factory, cur, prev)
    assert rerr <= 1e-13 and jerr <= RTOL, (method, rerr, jerr)


def test_options_no_gravity_upwind_weight(engine_factory):
    spec = problems.twop_lens((12, 10, 8), law="bc")
    spec.options.enable_gravity = False
    spec.options.upwind_weight = 0.75
    prev = _perturbed(spec, 4)
    cur = _perturbed(spec, 5)
    rerr, jerr, _ = _compare(spec, engine_factory, cur, prev)
    assert rerr <= 1e-13 and jerr <= RTOL, (rerr, jerr)


def test_residual_only_matches(engine_factory):
    spec = problems.twop_lens((20, 12), law="vg")
    cur, prev = _perturbed(spec, 8), _perturbed(spec, 9)
    e = engine_factory(spec)
    r1, _ = e.assemble(cur, prev, jacobian=True)
    r2, j2 = e.assemble(cur, prev, jacobian=False)
    assert j2 is None and np.array_equal(r1, r2)


def test_nonfinite_residual_is_reported(engine_factory):
    from dumux_b200.binding import DmxError
    spec = problems.twop_lens((8, 8), law="bc")
    cur = spec.initial.copy()
    cur[5, 0] = np.nan
    e = engine_factory(spec)
    with pytest.raises(DmxError):
        e.assemble(cur, spec.initial)


@pytest.mark.parametrize("cells,law", [((33, 9, 37), "bc"), ((70, 19, 21), "vg")])
def test_tile_kernel_chunks_bit_identical(engine_factory, cells, law):
    """Partial 32x8 tiles, several z-chunks per tile column and chunk-boundary halo layers of the fused tile kernel; the
    kernel executes the oracle's IEEE operation sequence (shared-log pow, div_by are exact), so equality is bitwise."""
    spec = problems.twop_lens(cells, law=law, heterogeneity_sigma=0.5)
    prev = _perturbed(spec, 11)
    cur = _perturbed(spec, 12)
    rerr, jerr, (res_o, jac_o, res_g, jac_g) = _compare(spec, engine_factory, cur, prev)
    assert rerr <= 1e-13 and jerr <= RTOL, (rerr, jerr)
    assert np.array_equal(res_g, res_o)
    assert np.array_equal(jac_g, jac_o), int((jac_g != jac_o).sum())


def test_one_dimensional_grid(engine_factory):
    """dim = 1 instantiation of the tile kernel and the structured solver path (YaspGrid<1>): assembly bit-identical, the
    stationary solve reproduces the oracle."""
    spec = problems.onep_incompressible((40,))
    cur = _perturbed(spec, 6, dp=1e4)
    rerr, jerr, (res_o, jac_o, res_g, jac_g) = _compare(spec, engine_factory, cur, None)
    assert np.array_equal(res_g, res_o) and np.array_equal(jac_g, jac_o)
    o = Oracle(spec)
    xo, sto, ito, redo = o.solve(jac_o, res_o, reduction=1e-12)
    e = engine_factory(spec)
    xg, stg, itg, redg = e.solve(jac_o, res_o, reduction=1e-12)
    assert sto == 0 and stg == 0 and np.linalg.norm(xg - xo) <= 1e-10 * np.linalg.norm(xo)
