"""N > 1 path on CPU: slab decomposition with overlap 1, owner-masked scalar products, copyOwnerToAll halo exchange and
per-rank ILU0 (overlapping Schwarz), run as 2 real processes over torch.distributed/gloo and compared with the
in-process multi-rank reference and with the single-domain solve.  The GPU path implements the same
sequence with NCCL (dumux_b200/csrc/dist.cu); tests/test_gpu_dist.py compares it with the same reference."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dumux_b200 import problems
from oracle import dist_oracle as D
from oracle import oracle_py as O

CELLS = (12, 10, 14)


def _make_spec(box):
    return problems.twop_lens(CELLS, law="bc", heterogeneity_sigma=0.4, box=box, plane_rng=True)


def _perturb(spec, box, seed=3):
    """deterministic global perturbation, cut to the local box"""
    n = int(np.prod(CELLS))
    rng = np.random.RandomState(seed)
    sn = rng.uniform(0.0, 0.25, size=n)
    dp = rng.uniform(-40.0, 40.0, size=n)
    if box is not None:
        sl = tuple(slice(*box[a]) for a in (2, 1, 0))
        sn = sn.reshape(CELLS[2], CELLS[1], CELLS[0])[sl].reshape(-1)
        dp = dp.reshape(CELLS[2], CELLS[1], CELLS[0])[sl].reshape(-1)
    u = spec.initial.copy()
    u[:, 0] += dp
    u[:, 1] = sn
    return u


def _rank_job(rank_obj):
    """what every rank does: assemble its slab, solve to a tight tolerance, take one Newton solve"""
    box = rank_obj.box if rank_obj.comm.nranks > 1 else None
    cur = _perturb(rank_obj.spec, box).reshape(-1)
    prev = rank_obj.spec.initial.reshape(-1)
    res, jac = rank_obj.o.assemble(cur, prev)
    x, st, its, red = rank_obj.bicgstab(jac, res, reduction=1e-11, maxit=500)
    xg, stg, itsg, redg = rank_obj.gmres(jac, res, reduction=1e-11, maxit=500, restart=10)
    u, nst, nsteps, lin_its = rank_obj.newton(rank_obj.spec.initial, rank_obj.spec.initial)
    xa, sta, itsa, reda = rank_obj.bicgstab(jac, res, reduction=1e-11, maxit=500, precond="amg")
    return {"x": x, "st": st, "its": its, "red": red, "res": res, "u": u, "nst": nst, "nsteps": nsteps, "lin_its": lin_its,
            "owner": rank_obj.owner.copy(), "xg": xg, "stg": stg, "itsg": itsg, "redg": redg, "jac": jac,
            "xa": xa, "sta": sta, "itsa": itsa}


def _gloo_worker(rank, world, port, q, part=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = _rank_job(D.BoxRank(_make_spec, CELLS, D.TorchComm(), part))
        q.put((rank, out))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def reference_runs():
    single = D.run_threads(_make_spec, CELLS, 1, _rank_job)[0]
    two = D.run_threads(_make_spec, CELLS, 2, _rank_job)
    three = D.run_threads(_make_spec, CELLS, 3, _rank_job)
    return single, two, three


def test_partition_layout():
    """overlap 1: rank r holds its owned layers plus one layer of each neighbour; every layer has exactly one owner"""
    for P in (2, 3, 4):
        owners = np.zeros(CELLS[2], dtype=int)
        for r in range(P):
            lo, hi, b0, b1 = problems.slab_partition(CELLS[2], P, r)
            assert lo == max(0, b0 - 1) and hi == min(CELLS[2], b1 + 1)
            owners[b0:b1] += 1
        assert np.all(owners == 1)


def test_overlap_rows_are_incomplete_but_owned_rows_match(reference_runs):
    """Residual rows of owned cells equal the single-domain rows bit for bit; the overlap rows lack the outer-face flux."""
    single, two, _ = reference_runs
    nf = CELLS[0] * CELLS[1] * 2
    glob = single["res"].reshape(CELLS[2], nf)
    for r, out in enumerate(two):
        lo, hi, b0, b1 = problems.slab_partition(CELLS[2], 2, r)
        loc = out["res"].reshape(hi - lo, nf)
        assert np.array_equal(loc[b0 - lo:b1 - lo], glob[b0:b1])
        ghost = [k for k in range(lo, hi) if not (b0 <= k < b1)]
        assert ghost and any(not np.array_equal(loc[k - lo], glob[k]) for k in ghost)


@pytest.mark.parametrize("P", [2, 3])
def test_schwarz_bicgstab_converges_to_single_domain_solution(reference_runs, P):
    single, two, three = reference_runs
    runs = two if P == 2 else three
    assert single["st"] == 0 and all(o["st"] == 0 for o in runs)
    x = D.gather_owned([o["x"] for o in runs], CELLS, P, 2)
    assert np.linalg.norm(x - single["x"]) <= 1e-8 * np.linalg.norm(single["x"])
    # per-rank ILU0 is a weaker preconditioner than the global one: more iterations, the same for every rank
    assert len({o["its"] for o in runs}) == 1 and runs[0]["its"] >= single["its"]
    # after copyOwnerToAll the overlap copies agree with the owner's values
    nf = CELLS[0] * CELLS[1] * 2
    for r in range(P - 1):
        lo0, hi0, b00, b10 = problems.slab_partition(CELLS[2], P, r)
        lo1, hi1, b01, b11 = problems.slab_partition(CELLS[2], P, r + 1)
        a = runs[r]["x"].reshape(hi0 - lo0, nf)
        b = runs[r + 1]["x"].reshape(hi1 - lo1, nf)
        assert np.array_equal(a[b10 - lo0], b[b10 - lo1]) and np.array_equal(a[b10 - 1 - lo0], b[b10 - 1 - lo1])


@pytest.mark.parametrize("P", [2, 3])
def test_schwarz_gmres_converges_to_single_domain_solution(reference_runs, P):
    """ILURestartedGMResIstlSolver on the overlapping decomposition: the single-rank run equals the sequential C++ restatement, the
    P-rank runs converge to the same global solution with overlap copies consistent."""
    single, two, three = reference_runs
    runs = two if P == 2 else three
    assert single["stg"] == 0 and all(o["stg"] == 0 for o in runs)
    o1 = O.Oracle(_make_spec(None))
    xs, sts, itss, reds = o1.solve_gmres(single["jac"], single["res"], reduction=1e-11, maxit=500, restart=10)
    assert sts == 0 and itss == single["itsg"] and np.linalg.norm(xs - single["xg"]) <= 1e-12 * np.linalg.norm(xs)
    x = D.gather_owned([o["xg"] for o in runs], CELLS, P, 2)
    assert np.linalg.norm(x - single["x"]) <= 1e-7 * np.linalg.norm(single["x"])
    assert len({o["itsg"] for o in runs}) == 1 and runs[0]["itsg"] >= single["itsg"]
    nf = CELLS[0] * CELLS[1] * 2
    for r in range(P - 1):
        lo0, hi0, b00, b10 = problems.slab_partition(CELLS[2], P, r)
        lo1, hi1, b01, b11 = problems.slab_partition(CELLS[2], P, r + 1)
        a = runs[r]["xg"].reshape(hi0 - lo0, nf)
        b = runs[r + 1]["xg"].reshape(hi1 - lo1, nf)
        assert np.allclose(a[b10 - lo0], b[b10 - lo1], rtol=1e-12, atol=0) and np.allclose(a[b10 - 1 - lo0], b[b10 - 1 - lo1], rtol=1e-12, atol=0)


@pytest.mark.parametrize("P", [2, 3])
def test_newton_iteration_count_independent_of_partition(reference_runs, P):
    """north_star: same Newton iteration count, fields to 1e-8 relative L2 (BiCGSTAB counts may depend on P)."""
    single, two, three = reference_runs
    runs = two if P == 2 else three
    assert single["nst"] == 0 and all(o["nst"] == 0 for o in runs)
    assert all(o["nsteps"] == single["nsteps"] for o in runs)
    u = D.gather_owned([o["u"] for o in runs], CELLS, P, 2).reshape(-1, 2)
    us = single["u"].reshape(-1, 2)
    assert np.linalg.norm(u[:, 0] - us[:, 0]) <= 1e-8 * np.linalg.norm(us[:, 0])
    assert np.linalg.norm(u[:, 1] - us[:, 1]) <= 1e-8 * max(1.0, np.linalg.norm(us[:, 1]))


def test_two_processes_over_gloo_match_reference(reference_runs):
    """world_size 2 over torch.distributed/gloo: identical to the in-process reference (same arithmetic, real messages)."""
    import torch.multiprocessing as mp
    _, two, _ = reference_runs
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert got[r]["st"] == 0 and got[r]["its"] == two[r]["its"]
        assert np.array_equal(got[r]["x"], two[r]["x"])
        assert got[r]["stg"] == 0 and got[r]["itsg"] == two[r]["itsg"] and np.array_equal(got[r]["xg"], two[r]["xg"])
        assert got[r]["nsteps"] == two[r]["nsteps"] and got[r]["lin_its"] == two[r]["lin_its"]
        assert np.array_equal(got[r]["u"], two[r]["u"])
        assert got[r]["sta"] == 0 and got[r]["itsa"] == two[r]["itsa"] and np.array_equal(got[r]["xa"], two[r]["xa"])


# ------------------------------------------------------------------------------------------------------------------------
# general Grid.Partitioning "px py pz" (io/grid/gridmanager_yasp.hh:194-203): blocks with face, edge and corner neighbours
# ------------------------------------------------------------------------------------------------------------------------
BLOCK_PARTS = [(2, 1, 2), (2, 2, 2), (1, 3, 2)]


def test_block_partition_layout():
    """every cell has exactly one owner; a block holds its owned range plus one layer towards every neighbour; ranks are
    numbered x fastest"""
    for part in BLOCK_PARTS + [(3, 2, 1)]:
        P = int(np.prod(part))
        owners = np.zeros(CELLS[::-1], dtype=int)
        for r in range(P):
            rng = problems.box_partition(CELLS, part, r)
            coord = problems.rank_coord(part, r)
            assert r == coord[0] + part[0] * (coord[1] + part[1] * coord[2])
            for a in range(3):
                lo, hi, b0, b1 = rng[a]
                assert lo == max(0, b0 - 1) and hi == min(CELLS[a], b1 + 1)
                assert (b1 - b0) in (CELLS[a] // part[a], CELLS[a] // part[a] + 1)
            owners[tuple(slice(rng[a][2], rng[a][3]) for a in (2, 1, 0))] += 1
        assert np.all(owners == 1)


@pytest.fixture(scope="module")
def block_runs(reference_runs):
    return {part: D.run_threads(_make_spec, CELLS, int(np.prod(part)), _rank_job, part) for part in BLOCK_PARTS}


@pytest.mark.parametrize("part", BLOCK_PARTS)
def test_block_decomposition_owned_rows_and_solution(reference_runs, block_runs, part):
    """Owned residual rows equal the single-domain rows bit for bit; Schwarz-BiCGSTAB / GMRes converge to the single-domain
    solution; after copyOwnerToAll every overlap copy (faces, edges, corners) equals its owner's value; same Newton count."""
    single = reference_runs[0]
    runs = block_runs[part]
    P = int(np.prod(part))
    assert np.array_equal(D.gather_owned([o["res"] for o in runs], CELLS, P, 2, part), single["res"])
    assert single["st"] == 0 and all(o["st"] == 0 for o in runs)
    x = D.gather_owned([o["x"] for o in runs], CELLS, P, 2, part)
    assert np.linalg.norm(x - single["x"]) <= 1e-8 * np.linalg.norm(single["x"])
    assert len({o["its"] for o in runs}) == 1 and runs[0]["its"] >= single["its"]
    xg = D.gather_owned([o["xg"] for o in runs], CELLS, P, 2, part)
    assert np.linalg.norm(xg - single["x"]) <= 1e-7 * np.linalg.norm(single["x"])
    # overlap copies == owner values: every rank's FULL local box equals the gathered global vector there
    glob = x.reshape(CELLS[2], CELLS[1], CELLS[0], 2)
    for r, o in enumerate(runs):
        rng = problems.box_partition(CELLS, part, r)
        sl = tuple(slice(rng[a][0], rng[a][1]) for a in (2, 1, 0))
        assert np.array_equal(o["x"].reshape(glob[sl].shape), glob[sl])
    assert all(o["nst"] == 0 and o["nsteps"] == single["nsteps"] for o in runs)
    u = D.gather_owned([o["u"] for o in runs], CELLS, P, 2, part).reshape(-1, 2)
    us = single["u"].reshape(-1, 2)
    assert np.linalg.norm(u[:, 0] - us[:, 0]) <= 1e-8 * np.linalg.norm(us[:, 0])
    assert np.linalg.norm(u[:, 1] - us[:, 1]) <= 1e-8 * max(1.0, np.linalg.norm(us[:, 1]))


def test_block_decomposed_amg_is_partition_independent(reference_runs, block_runs):
    """AMGBiCGSTABIstlSolver on a decomposition = the GLOBAL hierarchy cut like the grid (aggregates do not cross processor
    boundaries, smoother = BlockPreconditioner<SeqSSOR>): the iteration count stays at the single-domain count (+-1) whatever the
    partitioning, far below Schwarz-ILU0's; same solution; overlap copies equal their owners."""
    single, two, three = reference_runs
    assert single["sta"] == 0 and single["itsa"] * 3 < single["its"]
    cases = [(None, 2, two), (None, 3, three)] + [(part, int(np.prod(part)), block_runs[part]) for part in BLOCK_PARTS]
    for part, P, runs in cases:
        assert all(o["sta"] == 0 for o in runs) and len({o["itsa"] for o in runs}) == 1
        assert abs(runs[0]["itsa"] - single["itsa"]) <= 1, (part, runs[0]["itsa"], single["itsa"])
        x = D.gather_owned([o["xa"] for o in runs], CELLS, P, 2, part)
        assert np.linalg.norm(x - single["x"]) <= 1e-8 * np.linalg.norm(single["x"])
        glob = x.reshape(CELLS[2], CELLS[1], CELLS[0], 2)
        for r, o in enumerate(runs):
            rng = problems.box_partition(CELLS, part if part is not None else problems.default_partitioning(3, P), r)
            sl = tuple(slice(rng[a][0], rng[a][1]) for a in (2, 1, 0))
            assert np.array_equal(o["xa"].reshape(glob[sl].shape), glob[sl])


def test_amg_hierarchy_of_a_decomposition_tiles_the_global_levels():
    """level boxes: on every level the owned ranges of the ranks tile the global box of that level, each rank holds its owned
    range plus one overlap layer towards every neighbour, and an odd owned length ends in a single-cell aggregate"""
    from oracle.amg_oracle import AmgOracle
    cells, part = (13, 10, 9), (2, 1, 3)

    def make(box):
        return problems.twop_lens(cells, law="bc", heterogeneity_sigma=0.3, box=box)

    def job(ro):
        u0 = ro.spec.initial.reshape(-1)
        res, jac = ro.o.assemble(u0, u0)
        amg = AmgOracle(ro.local.cells, 3, ro.b, ro.o.rowptr, ro.o.colidx, jac, layout=ro.layout, part3=ro.part3, gcells=ro.cells)
        return [(lv.gcells, lv.ranges) for lv in amg.levels]

    out = D.run_threads(make, cells, 6, job, part)
    nlev = len(out[0])
    assert nlev >= 3 and all(len(o) == nlev for o in out)
    for l in range(nlev):
        g = out[0][l][0]
        owners = np.zeros(g[::-1], dtype=int)
        for r in range(6):
            gc, rng = out[r][l]
            assert gc == g
            for a in range(3):
                lo, hi, b0, b1 = rng[a]
                assert lo == max(0, b0 - 1) and hi == min(g[a], b1 + 1) and b1 > b0
            owners[tuple(slice(rng[a][2], rng[a][3]) for a in (2, 1, 0))] += 1
        assert np.all(owners == 1)
    assert int(np.prod(out[0][-1][0])) <= 8 or out[0][-1][0] == part


@pytest.mark.parametrize("part", [None] + BLOCK_PARTS)
def test_native_multi_rank_solver_equals_the_python_ranks(reference_runs, block_runs, part):
    """oracle.cpp orc_schwarz_ilu0_bicgstab (what bench.py's CPU arm times: one OpenMP thread per rank) against
    BoxRank.bicgstab rank by rank: same iteration count (+-1: the C dot sums the owned entries sequentially, numpy's in blocks),
    same solution incl. the overlap copies"""
    from oracle.amg_oracle import grid_pattern
    runs = reference_runs[1] if part is None else block_runs[part]
    part_ = part if part is not None else problems.default_partitioning(3, 2)
    P = int(np.prod(part_))
    systems = []
    for r in range(P):
        rng = problems.box_partition(CELLS, part_, r)
        rp, ci = grid_pattern(tuple(x[1] - x[0] for x in rng), 3)
        systems.append((rp, ci, runs[r]["jac"], runs[r]["res"]))
    xs, st, its, red, secs = D.native_schwarz_bicgstab(CELLS, part_, 2, systems, 1e-11, 500)
    assert st == 0 and abs(its - runs[0]["its"]) <= 1 and red <= 1e-11 and secs > 0
    for r in range(P):
        assert np.linalg.norm(xs[r] - runs[r]["x"]) <= 1e-8 * np.linalg.norm(runs[r]["x"])


def test_four_processes_over_gloo_block_partition(block_runs):
    """world_size 4 over gloo with Grid.Partitioning "2 1 2": identical to the in-process reference (real messages to face and
    edge neighbours)."""
    import torch.multiprocessing as mp
    part = (2, 1, 2)
    ref = block_runs[part]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29870 + os.getpid() % 100
    procs = [ctx.Process(target=_gloo_worker, args=(r, 4, port, q, part)) for r in range(4)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(4))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # (gloo sums the four contributions of an all-reduce in its own order, so the iterates agree to rounding, not bit for bit)
    for r in range(4):
        assert got[r]["st"] == 0 and got[r]["its"] == ref[r]["its"]
        assert np.linalg.norm(got[r]["x"] - ref[r]["x"]) <= 1e-9 * np.linalg.norm(ref[r]["x"])
        assert got[r]["nsteps"] == ref[r]["nsteps"] and all(abs(a - b) <= 1 for a, b in zip(got[r]["lin_its"], ref[r]["lin_its"]))
        assert np.linalg.norm(got[r]["u"] - ref[r]["u"]) <= 1e-8 * np.linalg.norm(ref[r]["u"])
