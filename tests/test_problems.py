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
