"""VTU output / restart: layout of the reference's VtkOutputModule files, loadSolution round trip."""
import os

import numpy as np
import pytest

from dumux_b200 import problems, vtkio
from oracle.oracle_py import Oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES_2P = ["S_aq", "p_aq", "rho_aq", "mob_aq", "S_napl", "p_napl", "rho_napl", "mob_napl", "pc", "porosity"]
ORACLE_COL = {"S_aq": 0, "S_napl": 1, "p_aq": 2, "p_napl": 3, "rho_aq": 4, "rho_napl": 5, "mob_aq": 6, "mob_napl": 7, "pc": 8, "porosity": 9}


def _fields_from_oracle(o, u):
    vv = o.volvars(u)
    return {nm: vv[:, ORACLE_COL[nm]] for nm in NAMES_2P}


def test_vtu_layout_matches_reference_files(tmp_path):
    """Same cell-data arrays, in the same order, as test_2p_incompressible_cc-reference.vtu (TwoPIOFields + process rank); values
    survive the Float32 ASCII round trip; points and connectivity describe the 48x32 quadrilateral grid."""
    import json
    spec = problems.twop_lens((48, 32), law="vg")
    o = Oracle(spec)
    u = spec.initial.reshape(-1).copy()
    fields = _fields_from_oracle(o, u)
    path = str(tmp_path / "lens-00000.vtu")
    vtkio.write_vtu(path, problems.node_coords(spec.cells, spec.lower, spec.upper), fields)
    n, back = vtkio.read_vtu(path)
    assert n == 1536 and list(back) == NAMES_2P + ["process rank"]
    manifest = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))["test_2p_incompressible_cc"]
    assert sorted(back) == manifest["fields"] and n == manifest["cells"]
    for nm in NAMES_2P:
        assert np.allclose(back[nm], fields[nm], rtol=6e-6, atol=0)        # six significant digits, like the reference files
    import xml.etree.ElementTree as ET
    piece = ET.parse(path).getroot().find("UnstructuredGrid/Piece")
    assert piece.get("NumberOfPoints") == "1617"                       # 49 x 33, as in the reference file
    pts = np.array(piece.find("Points/DataArray").text.split(), dtype=float).reshape(-1, 3)
    assert pts[:, 0].max() == 6.0 and pts[:, 1].max() == 4.0 and not pts[:, 2].any()
    conn = np.array(piece.find("Cells/DataArray[@Name='connectivity']").text.split(), dtype=int).reshape(-1, 4)
    # first cell: vertices 0, 1, 50, 49 (VTK quad order over the lexicographic 49-wide node grid)
    assert list(conn[0]) == [0, 1, 50, 49]
    c = pts[conn].mean(axis=1)[:, :2]
    assert np.allclose(c, problems.cell_centers(spec.cells, spec.lower, spec.upper), atol=1e-6)


def test_restart_from_vtu_continues_the_run(tmp_path):
    """loadSolution semantics: a run written at t = 1000 s and restarted from the file ends where the uninterrupted run ends, up to
    the Float32 precision of the file (the reference's test_2p_incompressible_tpfa_restart does the same)."""
    spec = problems.twop_lens((24, 16), law="vg")
    o = Oracle(spec)
    u_mid, n1, its1, dts1 = o.run_timeloop(spec.initial, 1000.0, 250.0)
    path = str(tmp_path / "restart.vtu")
    vtkio.write_vtu(path, problems.node_coords(spec.cells, spec.lower, spec.upper), _fields_from_oracle(o, u_mid))
    u0 = vtkio.load_solution(path, ["p_aq", "S_napl"])
    um = u_mid.reshape(-1, 2)
    assert np.abs(u0[:, 0] / um[:, 0] - 1).max() <= 6e-6 and np.abs(u0[:, 1] - um[:, 1]).max() <= 6e-6      # six significant digits
    u_a, *_ = Oracle(spec).run_timeloop(u0, 500.0, dts1[-1])
    u_b, *_ = Oracle(spec).run_timeloop(u_mid, 500.0, dts1[-1])
    a, b = u_a.reshape(-1, 2), u_b.reshape(-1, 2)
    assert np.abs(a[:, 1] - b[:, 1]).max() < 1e-5 and np.abs(a[:, 0] / b[:, 0] - 1).max() < 1e-6
    with pytest.raises(KeyError):
        vtkio.load_solution(path, ["p_liq"])


def test_2p_incompressible_tpfa_restart_as_the_reference_runs_it(tmp_path):
    """test_2p_incompressible_tpfa_restart (test/porousmediumflow/2p/incompressible/CMakeLists.txt:30-40): the reference restarts
    the 48 x 32 lens run from its fifth output file with `-Restart.Time 2054.01 -TimeLoop.DtInitial 603.14` and compares the SECOND
    output of the restarted run with test_2p_incompressible_cc-reference.vtu (t = 3000 s).  Two things are pinned by that command
    line: (i) the uninterrupted run is at t = 2054.01 s after five time steps -- the step sizes follow from the Newton iteration
    counts (suggestTimeStepSize), so the counts of the first four steps are the reference's; (ii) from the restart file (Float32
    ASCII like the reference's) two more time steps reach t = 3000 s with the golden fields."""
    spec = problems.twop_lens((48, 32), law="vg")
    o = Oracle(spec)
    u_all, n_all, its_all, dts_all = o.run_timeloop(spec.initial, 3000.0, 250.0)
    assert n_all == 7 and abs(np.sum(dts_all[:5]) - 2054.01) <= 5e-3          # "Restart.Time 2054.01" (six digits)
    u5, n5, its5, dts5 = o.run_timeloop(spec.initial, float(np.sum(dts_all[:5])), 250.0)
    assert n5 == 5 and list(its5) == list(its_all[:5])
    path = str(tmp_path / "test_2p_incompressible_tpfa-00005.vtu")
    vtkio.write_vtu(path, problems.node_coords(spec.cells, spec.lower, spec.upper), _fields_from_oracle(o, u5))
    u0 = vtkio.load_solution(path, ["p_aq", "S_napl"])
    u_end, n2, its2, dts2 = Oracle(spec).run_timeloop(u0, 3000.0 - 2054.01, 603.14)
    assert n2 == 2 and dts2[0] == 603.14                                       # the compared file is output number 00002
    g = np.load(os.path.join(GOLDEN, "test_2p_incompressible_cc.npz"))
    for name, col in (("p_aq", 0), ("S_napl", 1)):
        ref = g[name].astype(np.float64)
        d = np.abs(u_end.reshape(-1, 2)[:, col] - ref)
        assert np.all((d <= 1.5e-7) | (d <= 1e-2 * np.abs(ref))), name          # the reference's fuzzy bar


def test_3d_and_1d_grids(tmp_path):
    for cells in ((4, 3, 2), (5,)):
        dim = len(cells)
        coords = problems.node_coords(cells, (0.0,) * dim, (1.0,) * dim)
        n = int(np.prod(cells))
        path = str(tmp_path / f"g{dim}.vtu")
        vtkio.write_vtu(path, coords, {"p": np.arange(n, dtype=float)})
        m, back = vtkio.read_vtu(path)
        assert m == n and np.array_equal(back["p"], np.arange(n, dtype=np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("oilwet", [False, True])
def test_device_output_fields_equal_oracle_volvars(engine_factory, oilwet, tmp_path):
    """dmx_output_fields: TwoPIOFields of the device-resident state, bit-identical to the oracle's volume variables (incl. the
    oil-wet lens); written as VTU and compared with the golden file of the corresponding reference test at its fuzzy bar."""
    from dumux_b200 import binding as B
    spec = problems.twop_lens((48, 32), law="vg", oilwet=oilwet, dt=130.0 if oilwet else 250.0)
    o = Oracle(spec)
    e = engine_factory(spec)
    for eng in (o, e):
        eng.set_linear_solver("gmres", 10)
    ug, its, dts = e.run_timeloop(spec.initial, 3000.0, spec.options.dt)
    fields = e.output_fields()
    vv = o.volvars(ug)
    assert list(fields) == NAMES_2P
    for nm in NAMES_2P:
        assert np.array_equal(fields[nm], vv[:, ORACLE_COL[nm]]), nm
    path = str(tmp_path / "out.vtu")
    vtkio.write_vtu(path, problems.node_coords(spec.cells, spec.lower, spec.upper), fields)
    n, back = vtkio.read_vtu(path)
    g = np.load(os.path.join(GOLDEN, "test_2p_incompressible_tpfa_oilwet.npz" if oilwet else "test_2p_incompressible_cc.npz"))
    for nm in NAMES_2P:
        a, ref = back[nm].astype(np.float64), g[nm].astype(np.float64)
        d = np.abs(a - ref)
        assert np.all((d <= 1.5e-7) | (d <= 1e-2 * np.maximum(np.abs(a), np.abs(ref)))), nm


def test_parallel_pieces_and_pvtu(tmp_path):
    """Per-rank pieces hold the OWNED cells only (as the reference's s0002-p0000-*.vtu pieces do: the rank-0 piece of
    test_richards_lens_tpfa_parallel-reference.vtu has half of the 24x16 cells), carry `process rank`, follow Dune's parallel file
    names, and together reproduce the single-domain field; a rank restarts from its own piece (loadSolution) with the overlap
    left for the owner -> copy communication."""
    cells, part = (9, 7, 6), (2, 1, 2)
    P = int(np.prod(part))
    coords = problems.node_coords(cells, (0.0, 0.0, 0.0), (3.0, 2.0, 1.0))
    n = int(np.prod(cells))
    rng = np.random.RandomState(4)
    # values with at most six significant digits: the ASCII writer prints %.6g like Dune's
    glob = {"p_aq": (1e5 + rng.randint(0, 1000, size=n)).astype(np.float32).astype(np.float64),
            "S_napl": np.round(rng.uniform(0, 0.5, size=n), 3).astype(np.float32).astype(np.float64)}
    g3 = {k: v.reshape(cells[::-1]) for k, v in glob.items()}
    merged = {k: np.full(cells[::-1], np.nan) for k in glob}
    for r in range(P):
        rngs = problems.box_partition(cells, part, r)
        box = tuple(slice(rngs[a][0], rngs[a][1]) for a in (2, 1, 0))
        local = {k: v[box].reshape(-1) for k, v in g3.items()}
        nodes = [coords[a][rngs[a][0]:rngs[a][1] + 1] for a in range(3)]
        own_lo = [rngs[a][2] - rngs[a][0] for a in range(3)]
        own_hi = [rngs[a][3] - rngs[a][0] for a in range(3)]
        path = vtkio.write_piece(str(tmp_path), "run", 3, P, r, nodes, own_lo, own_hi, local)
        assert os.path.basename(path) == f"s{P:04d}-p{r:04d}-run-00003.vtu"
        m, back = vtkio.read_vtu(path)
        assert m == int(np.prod([rngs[a][3] - rngs[a][2] for a in range(3)]))
        assert np.all(back["process rank"] == r)
        own = tuple(slice(rngs[a][2], rngs[a][3]) for a in (2, 1, 0))
        for k in glob:
            merged[k][own] = back[k].reshape(merged[k][own].shape)
    master = vtkio.write_pvtu(str(tmp_path), "run", 3, P, {"p_aq": 1, "S_napl": 1})
    assert os.path.basename(master) == f"s{P:04d}-run-00003.pvtu"
    pieces, names = vtkio.read_pvtu(master)
    assert len(pieces) == P and names == ["p_aq", "S_napl", "process rank"] and all(os.path.exists(p) for p in pieces)
    for k in glob:
        assert np.array_equal(merged[k].reshape(-1), glob[k])
    # restart of rank 1: owned cells from its piece, overlap cells untouched (zero) until the halo exchange
    r = 1
    rngs = problems.box_partition(cells, part, r)
    lc = [rngs[a][1] - rngs[a][0] for a in range(3)]
    own_lo = [rngs[a][2] - rngs[a][0] for a in range(3)]
    own_hi = [rngs[a][3] - rngs[a][0] for a in range(3)]
    u = vtkio.load_solution_piece(master, r, ["p_aq", "S_napl"], lc, own_lo, own_hi).reshape(lc[2], lc[1], lc[0], 2)
    own = tuple(slice(own_lo[a], own_hi[a]) for a in (2, 1, 0))
    gl = tuple(slice(rngs[a][2], rngs[a][3]) for a in (2, 1, 0))
    assert np.array_equal(u[own][..., 0], g3["p_aq"][gl]) and np.array_equal(u[own][..., 1], g3["S_napl"][gl])
    mask = np.ones(u.shape[:3], dtype=bool)
    mask[own] = False
    assert mask.any() and np.all(u[mask] == 0.0)
