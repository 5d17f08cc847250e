"""Host-side set-up logic: flat-array sampling of Problem/SpatialParams, slab partitioning (Grid.Partitioning "1 1 P",
Grid.Overlap 1) and the slab-local generation the multi-GPU bench uses."""
import numpy as np
import pytest

from dumux_b200 import problems as P


def test_cell_and_face_ordering():
    ctr = P.cell_centers((3, 2), (0.0, 0.0), (3.0, 2.0))
    assert ctr.shape == (6, 2)
    assert np.allclose(ctr[:4], [[0.5, 0.5], [1.5, 0.5], [2.5, 0.5], [0.5, 1.5]])      # x fastest (YaspGrid)
    fc = P.side_face_centers((3, 2, 4), (0, 0, 0), (3.0, 2.0, 4.0), 1)                  # +x side: y fastest, then z
    assert fc.shape == (8, 3) and np.allclose(fc[:3], [[3.0, 0.5, 0.5], [3.0, 1.5, 0.5], [3.0, 0.5, 1.5]])


@pytest.mark.parametrize("N,Pn", [(10, 3), (256, 8), (7, 7), (512, 2), (5, 1)])
def test_slab_partition_covers_and_overlaps(N, Pn):
    owned = []
    for r in range(Pn):
        lo, hi, b0, b1 = P.slab_partition(N, Pn, r)
        assert lo == (max(0, b0 - 1) if Pn > 1 else 0) and hi == (min(N, b1 + 1) if Pn > 1 else N)
        owned.append((b0, b1))
    assert owned[0][0] == 0 and owned[-1][1] == N
    assert all(a[1] == b[0] for a, b in zip(owned, owned[1:]))
    sizes = [b - a for a, b in owned]
    assert max(sizes) - min(sizes) <= 1


def test_slab_local_spec_equals_cut_of_global():
    cells = (8, 6, 10)
    g = P.twop_lens(cells, law="bc", heterogeneity_sigma=0.3, plane_rng=True)
    for r in range(3):
        lo, hi, b0, b1 = P.slab_partition(10, 3, r)
        l = P.twop_lens(cells, law="bc", heterogeneity_sigma=0.3, slab=(lo, hi))
        assert np.array_equal(g.K.reshape(10, -1)[lo:hi].reshape(-1), l.K)
        assert np.array_equal(g.region.reshape(10, -1)[lo:hi].reshape(-1), l.region)
        assert np.array_equal(g.initial.reshape(10, -1, 2)[lo:hi].reshape(-1, 2), l.initial)
        for s in range(4):
            assert np.array_equal(g.bc_type[s].reshape(10, -1)[lo:hi].reshape(-1), l.bc_type[s])
            assert np.array_equal(g.bc_values[s].reshape(10, -1, 2)[lo:hi].reshape(-1, 2), l.bc_values[s])
        for s in (4, 5):
            assert np.array_equal(g.bc_type[s], l.bc_type[s])


def test_lens_problem_matches_reference_setup():
    """test/porousmediumflow/2p/incompressible: spatialparams.hh:46-140, problem.hh:50-168, params.input."""
    s = P.twop_lens((48, 32), law="vg")
    assert s.num_cells == 1536 and s.num_eq == 2
    assert sorted(set(s.K)) == [9.05e-12, 4.6e-10]
    assert s.region.sum() == 24 * 8                       # lens [1,4]x[2,3] on the 48x32 grid
    assert s.materials[0].params == (0.0037, 4.7, 0.5) and s.materials[1].params == (0.00045, 7.3, 0.5)
    assert (s.materials[0].swr, s.materials[1].swr) == (0.05, 0.18)
    assert s.rho == (1000.0, 1460.0) and s.mu == (1e-3, 5.7e-4)
    top = s.bc_values[3]
    inlet = top[:, 1] != 0
    assert inlet.sum() == 8 and np.all(top[inlet, 1] == -0.04)   # 2/3 > (6-x)/6 > 1/2  <=>  2 < x < 3
    assert np.all(s.bc_type[0] == P.BC_DIRICHLET) and np.all(s.bc_type[1] == P.BC_DIRICHLET)
    assert np.all(s.bc_type[2] == P.BC_NEUMANN)
    # hydrostatic initial condition: p = 1e5 + rho*g*depth
    ctr = P.cell_centers((48, 32), (0, 0), (6.0, 4.0))
    assert np.allclose(s.initial[:, 0], 1e5 + 1000.0 * 9.81 * (4.0 - ctr[:, 1]))


def test_mt19937_lognormal_is_deterministic_and_lens_aware():
    a = P.lognormal_permeability(100, 1e-10, 0)
    b = P.lognormal_permeability(100, 1e-10, 0)
    assert np.array_equal(a, b) and np.all(a > 0)
    # ln K ~ N(ln 1e-10, 0.1*|ln 1e-10|): sample statistics in the right ballpark
    assert abs(np.log(a).mean() - np.log(1e-10)) < 1.0 and 1.5 < np.log(a).std() < 3.2


def test_buckleyleverett_problem_matches_reference_setup():
    """test/porousmediumflow/2p/buckleyleverett: params.input, problem.hh:52-140, spatialparams.hh:40-100"""
    s = P.twop_buckleyleverett()
    assert s.cells == (100, 1) and s.upper == (100.0, 75.0) and not s.options.enable_gravity
    m = s.materials[0]
    assert m.law == P.LAW_BC and m.params == (0.0, 4.0) and (m.swr, m.snr) == (0.2, 0.2)
    assert np.all(s.K == 1.01936799e-14) and np.all(s.phi == 0.2) and s.rho == (1000.0, 1000.0) and s.mu == (1e-3, 1e-3)
    assert np.all(s.bc_type[0] == P.BC_DIRICHLET) and np.all(s.bc_values[0] == [2e5, 0.2])          # left: p = 2e5, Sn = Snr
    assert np.all(s.bc_type[1] == P.BC_NEUMANN) and np.allclose(s.bc_values[1], [0.0, 3e-4])        # right: v_t * rho_n leaves
    assert np.all(s.bc_type[2] == P.BC_NEUMANN) and not s.bc_values[2].any() and not s.bc_values[3].any()
    assert np.all(s.initial == [2e5, 0.8])


def test_pointsource_convergence_and_extrusion_problems():
    ps = P.onep_pointsource()
    hot = np.flatnonzero(ps.source[:, 0])
    ctr = P.cell_centers(ps.cells, ps.lower, ps.upper)
    assert hot.size == 4 and np.allclose(np.abs(ctr[hot]), 0.01)               # the four cells around the origin of [-1,1]^2
    assert ps.source[hot, 0].sum() * 0.02 * 0.02 == pytest.approx(10.0)       # 10 kg/s in total
    cv = P.onep_convergence((20, 20))
    assert cv.K.shape == (400, 2) and np.all(cv.K[:, 0] == 1.0) and np.allclose(cv.K[:, 1], np.exp(-2.0))
    assert all(np.all(cv.bc_type[s] == P.BC_DIRICHLET) for s in range(4))
    fc = P.side_face_centers(cv.cells, cv.lower, cv.upper, 3)
    assert np.allclose(cv.bc_values[3][:, 0], P.onep_convergence_exact(fc[:, 0], fc[:, 1]))
    ex = P.onep_extrude()
    assert ex.options.extrusion == 10.0 and not ex.options.enable_gravity and np.all(ex.K == 1e-10)


def test_scheidegger_normal_entry():
    """D = (aL - aT) v v^T / |v| + aT |v| I: along the flow n.D.n = aL |v|, across it aT |v|; zero velocity -> zero"""
    v = np.array([[2e-5, 0.0], [0.0, 3e-5], [0.0, 0.0], [3e-5, 4e-5]])
    d0 = P.scheidegger_normal_entry(v, 0, 0.02, 0.008)
    d1 = P.scheidegger_normal_entry(v, 1, 0.02, 0.008)
    assert np.allclose(d0[:3], [0.02 * 2e-5, 0.008 * 3e-5, 0.0]) and np.allclose(d1[:3], [0.008 * 2e-5, 0.02 * 3e-5, 0.0])
    assert d0[3] == pytest.approx(0.012 * 9e-10 / 5e-5 + 0.008 * 5e-5)
    ts = P.tracer_constvel((10, 8), implicit=True, alpha_l=0.02, alpha_t=0.008)
    # the entry of a face is the same seen from both cells
    nx = 10
    d = ts.tracer_dispersion.reshape(8, 10, 4)
    assert np.allclose(d[:, :-1, 1], d[:, 1:, 0]) and np.allclose(d[:-1, :, 3], d[1:, :, 2])
