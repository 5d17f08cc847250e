"""Multi-GPU parity (needs >= 2 visible GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`):
one process per GPU, slab decomposition with overlap 1, NCCL halo exchange + all-reduce (dumux_b200/csrc/dist.cu),
compared with the CPU multi-rank reference (oracle/dist_oracle.py) rank by rank."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

CELLS = (20, 18, 22)


def _make_spec(box):
    from dumux_b200 import problems
    return problems.twop_lens(CELLS, law="bc", heterogeneity_sigma=0.4, box=box, plane_rng=True)


def _perturb(spec, box, seed=3):
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


def _cpu_rank_job(rank_obj):
    box = rank_obj.box if rank_obj.comm.nranks > 1 else None
    cur = _perturb(rank_obj.spec, box).reshape(-1)
    prev = rank_obj.spec.initial.reshape(-1)
    res, jac = rank_obj.o.assemble(cur, prev)
    x, st, its, red = rank_obj.bicgstab(jac, res, reduction=1e-10, maxit=500)
    xg, stg, itsg, redg = rank_obj.gmres(jac, res, reduction=1e-10, maxit=500, restart=10)
    u, nst, nsteps, lin_its = rank_obj.newton(rank_obj.spec.initial, rank_obj.spec.initial)
    # BlockPreconditioner<SeqSSOR> (SSORBiCGSTABIstlSolver) and <ParMTSSOR> on the overlapping decomposition
    other = {pc: rank_obj.bicgstab(jac, res, reduction=1e-8, maxit=2000, precond=pc) for pc in ("ssor", "par_mt_ssor", "amg")}
    # one V-cycle of the block-decomposed GLOBAL AMG hierarchy (oracle/amg_oracle.py), and its coarse matrices
    from oracle.amg_oracle import AmgOracle
    amg = AmgOracle(rank_obj.local.cells, 3, rank_obj.b, rank_obj.o.rowptr, rank_obj.o.colidx, jac, layout=rank_obj.layout,
                    part3=rank_obj.part3, gcells=rank_obj.cells)
    other["amg_v"] = amg.apply(res)
    other["amg_levels"] = [lv.cells for lv in amg.levels]
    other["amg_mats"] = [lv.values for lv in amg.levels]
    return {"res": res, "jac": jac, "x": x, "st": st, "its": its, "u": u, "nst": nst, "nsteps": nsteps, "lin_its": lin_its,
            "xg": xg, "stg": stg, "itsg": itsg, "redg": redg, "other": other}


def _gpu_worker(rank, world, uid, q, part=None):
    try:
        from dumux_b200 import binding as B
        from dumux_b200 import problems
        part_ = part if part is not None else problems.default_partitioning(3, world)
        rng = problems.box_partition(CELLS, part_, rank)
        box = [(r[0], r[1]) for r in rng]
        spec = _make_spec(box)
        eng = B.Engine(spec, device=rank, nccl_uid=uid, rank=rank, nranks=world, part=part)
        assert [(int(eng.own_lo[a]), int(eng.own_hi[a])) for a in range(3)] == [(r[2] - r[0], r[3] - r[0]) for r in rng]
        cur = _perturb(spec, box)
        res, jac = eng.assemble(cur, spec.initial)
        x, st, its, red = eng.solve(jac, res, reduction=1e-10, maxit=500)
        eng.set_linear_solver("gmres", 10)
        xg, stg, itsg, redg = eng.solve(jac, res, reduction=1e-10, maxit=500)
        eng.set_linear_solver("bicgstab")
        other = {name: eng.solve(jac, res, reduction=1e-8, maxit=2000, precond=pc)
                 for name, pc in (("ssor", B.PRECOND_SSOR), ("par_mt_ssor", B.PRECOND_PARMT_SSOR), ("amg", B.PRECOND_AMG))}
        eng.upload_jacobian(jac)
        eng.upload(B.VEC_WORK0, res)
        eng.precond_apply(B.PRECOND_AMG, B.VEC_WORK0, B.VEC_WORK1)
        other["amg_v"] = eng.download(B.VEC_WORK1)
        other["amg_levels"] = eng.amg_levels()
        other["amg_mats"] = [None] + [eng.amg_level_matrix(l) for l in range(1, len(other["amg_levels"]))]
        # halo exchange primitive: fill a vector with the rank id, exchange: every cell then carries its OWNER's rank id
        v = np.full(eng.n * eng.b, float(rank))
        eng.upload(B.VEC_WORK1, v)
        eng.halo_exchange(B.VEC_WORK1)
        halo = eng.download(B.VEC_WORK1).reshape(-1, eng.b)[:, 0].copy()
        nrm = eng.norm(B.VEC_RESIDUAL)
        u, nst, rep = eng.newton(spec.initial, spec.initial)
        # parallel output + restart: my piece holds my OWNED cells; a restart reads it back and the overlap is filled from the
        # owners by copyOwnerToAll (io/vtkoutputmodule.hh:346, io/loadsolution.hh:43,332)
        import tempfile
        from dumux_b200 import vtkio
        outdir = tempfile.mkdtemp(prefix=f"dmx_pvtu_r{rank}_")
        u2 = u.reshape(-1, 2)
        nodes = [problems.node_coords(CELLS, spec.lower, spec.upper)[a][box[a][0]:box[a][1] + 1] for a in range(3)]
        vtkio.write_piece(outdir, "restart", 1, world, rank, nodes, eng.own_lo, eng.own_hi, {"p_aq": u2[:, 0], "S_napl": u2[:, 1]})
        master = vtkio.write_pvtu(outdir, "restart", 1, world, {"p_aq": 1, "S_napl": 1})
        pieces, _ = vtkio.read_pvtu(master)
        own = vtkio.load_solution(pieces[rank], ["p_aq", "S_napl"])
        lc = [int(c) for c in eng.local_cells]
        loc = np.zeros((lc[2], lc[1], lc[0], 2))
        loc[tuple(slice(int(eng.own_lo[a]), int(eng.own_hi[a])) for a in (2, 1, 0))] = own.reshape(
            tuple(int(eng.own_hi[a] - eng.own_lo[a]) for a in (2, 1, 0)) + (2,))
        eng.upload(B.VEC_WORK1, loc.reshape(-1))
        eng.halo_exchange(B.VEC_WORK1)
        restarted = eng.download(B.VEC_WORK1)
        stored = np.array([float("%.6g" % v) for v in u.astype(np.float32)])          # what a %.6g Float32 file keeps of u
        restart_ok = bool(np.array_equal(restarted.astype(np.float32), stored.astype(np.float32)))
        # cross-rank failure agreement (assembly/fvassembler.hh:504-509 comm.min): a NaN on ONE rank makes EVERY rank
        # return DMX_STATUS_NONFINITE from the assembly instead of leaving the others in the next collective
        bad = spec.initial.copy()
        if rank == world - 1:
            bad[eng.n // 2, 0] = np.nan
        eng.upload(B.VEC_CUR, bad)
        st_bad = eng.assemble_device(True)
        p = eng.newton_params()
        st_step = eng.newton_step(p)[0]
        # ... and the engine is usable afterwards (the dt-halving retry of NewtonSolver::solve)
        eng.upload(B.VEC_CUR, spec.initial)
        st_ok = eng.assemble_device(True)
        q.put((rank, {"res": res, "jac": jac, "x": x, "st": st, "its": its, "halo": halo, "norm": nrm, "u": u, "nst": nst,
                      "nsteps": rep.newton_iterations, "lin_its": [rep.linear_iterations[i] for i in range(rep.newton_iterations)],
                      "launches": eng.launches(), "xg": xg, "stg": stg, "itsg": itsg, "redg": redg,
                      "st_bad": st_bad, "st_step": st_step, "st_ok": st_ok, "other": other,
                      "restart_ok": restart_ok}))
        eng.close()
    except BaseException as e:      # noqa: BLE001
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))
        raise


@pytest.mark.parametrize("world,part", [(2, None), (2, (2, 1, 1)), (4, (2, 1, 2)), (8, (2, 2, 2))])
def test_decomposed_newton_step_matches_cpu_reference(world, part):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run under gpurun --gpus {world})")
    import torch.multiprocessing as mp
    from dumux_b200 import binding as B
    from dumux_b200 import problems
    from oracle import dist_oracle as D
    # the oracle adds every scalar product in the device's reduction tree and the per-rank sums in rank order (as dist.cu does)
    # -> identical Krylov iterations for any number of ranks
    ref = D.run_threads(_make_spec, CELLS, world, _cpu_rank_job, part, gpu_reduction=True)
    uid = B.Engine.nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, uid, q, part)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert "error" not in got[r], got[r].get("error")
    part_ = part if part is not None else problems.default_partitioning(3, world)
    # owner rank of every global cell
    owner_of = np.zeros(CELLS[::-1])
    for r in range(world):
        rng = problems.box_partition(CELLS, part_, r)
        owner_of[tuple(slice(rng[a][2], rng[a][3]) for a in (2, 1, 0))] = r
    for r in range(world):
        g, c = got[r], ref[r]
        rng = problems.box_partition(CELLS, part_, r)
        # assembly of the local box incl. overlap rows (no scvf on the processor boundary): bit-exact
        assert np.array_equal(g["res"], c["res"]) and np.array_equal(g["jac"], c["jac"])
        # halo: after copyOwnerToAll every local cell (faces, edges, corners of the overlap included) carries its owner's id
        assert np.array_equal(g["halo"], owner_of[tuple(slice(rng[a][0], rng[a][1]) for a in (2, 1, 0))].reshape(-1))
        # owner-masked, all-reduced norm is the same number on every rank
        assert g["norm"] == got[0]["norm"]
        # Schwarz-BiCGSTAB: same iteration count, solution at the solver tolerance
        assert g["st"] == 0 and c["st"] == 0 and g["its"] == c["its"], (g["its"], c["its"])
        assert np.linalg.norm(g["x"] - c["x"]) <= 1e-7 * np.linalg.norm(c["x"])
        # Schwarz-GMRes(10) (ILURestartedGMResIstlSolver on the overlapping decomposition): same count, reduction and solution
        assert g["stg"] == 0 and c["stg"] == 0 and g["itsg"] == c["itsg"], (g["itsg"], c["itsg"])
        if g["itsg"] == c["itsg"]:
            assert g["redg"] == pytest.approx(c["redg"], rel=1e-5)
        assert np.linalg.norm(g["xg"] - c["xg"]) <= 1e-7 * np.linalg.norm(c["xg"])
        assert np.linalg.norm(g["xg"] - g["x"]) <= 1e-6 * np.linalg.norm(g["x"])          # both solve the same global system
        # the other block preconditioners on the decomposition: same counts (2 ranks: identical iteration), same solution
        for name in ("ssor", "par_mt_ssor", "amg"):
            (xa, sta, ita, _), (xb, stb, itb, _) = g["other"][name], c["other"][name]
            assert sta == 0 and stb == 0 and ita == itb, (name, ita, itb)
            assert np.linalg.norm(xa - xb) <= 1e-6 * np.linalg.norm(xb)
        # AMG on the decomposition = the GLOBAL hierarchy cut like the grid: same level boxes, bit-identical Galerkin matrices
        # and V-cycle on every rank, and a mesh- and partition-independent iteration count (far below Schwarz-ILU0's)
        assert g["other"]["amg_levels"] == c["other"]["amg_levels"] and len(c["other"]["amg_levels"]) >= 3
        for l in range(1, len(c["other"]["amg_levels"])):
            assert np.array_equal(g["other"]["amg_mats"][l], c["other"]["amg_mats"][l]), (r, l)
        assert np.array_equal(g["other"]["amg_v"], c["other"]["amg_v"])
        assert g["other"]["amg"][2] * 3 < g["its"]
        # Newton: same iteration count, fields to 1e-8
        assert g["nst"] == 0 and g["nsteps"] == c["nsteps"]
        assert g["lin_its"] == c["lin_its"], (g["lin_its"], c["lin_its"])
        ug, uc = g["u"].reshape(-1, 2), c["u"].reshape(-1, 2)
        assert np.linalg.norm(ug[:, 0] - uc[:, 0]) <= 1e-8 * np.linalg.norm(uc[:, 0])
        assert np.linalg.norm(ug[:, 1] - uc[:, 1]) <= 1e-8 * max(1.0, np.linalg.norm(uc[:, 1]))
        assert g["launches"] > 0
        # failure agreement: every rank reports the NaN that only the last rank holds, and recovers
        assert g["st_bad"] == B.STATUS_NONFINITE and g["st_step"] == B.STATUS_NONFINITE and g["st_ok"] == 0
        assert g["restart_ok"]
    # global norm = sqrt(sum of owned squares)
    owned = D.gather_owned([ref[r]["res"] for r in range(world)], CELLS, world, 2, part)
    assert abs(got[0]["norm"] - np.linalg.norm(owned)) <= 1e-12 * np.linalg.norm(owned)


# ------------------------------------------------------------------------------------------------------------------------
# tracer transport (BASELINE config 5), slab-decomposed: explicit steps on 2 GPUs against the single-domain CPU oracle
# ------------------------------------------------------------------------------------------------------------------------
TRACER_CELLS = (18, 14, 20)
TRACER_STEPS = 12


def _tracer_problem():
    from dumux_b200 import problems
    from oracle.oracle_py import Oracle
    ps = problems.onep_tracer_pressure(TRACER_CELLS)
    n = int(np.prod(TRACER_CELLS))
    rng = np.random.RandomState(11)
    ctr = problems.cell_centers(TRACER_CELLS, ps.lower, ps.upper)
    p = 1.0e5 * (1.1 - 0.1 * ctr[:, 2]) + rng.uniform(-20.0, 20.0, size=n)       # mostly upward flow + noise
    vf = Oracle(ps).volume_flux(p)
    ts = problems.tracer_transport(TRACER_CELLS, vf, dt=0.01)      # explicit Euler: below the CFL limit of the noisy field
    ts.initial[:, 0] = rng.uniform(0.0, 2e-11, size=n)
    return ts


def _tracer_worker(rank, world, uid, q):
    try:
        from dumux_b200 import binding as B
        from dumux_b200 import problems
        ts = _tracer_problem()
        eng = B.Engine(ts, device=rank, nccl_uid=uid, rank=rank, nranks=world)
        lo, hi, b0, b1 = problems.slab_partition(TRACER_CELLS[2], world, rank)
        nf = TRACER_CELLS[0] * TRACER_CELLS[1]
        x0 = ts.initial.reshape(TRACER_CELLS[2], nf)[lo:hi].reshape(-1)
        eng.upload(B.VEC_CUR, x0)
        eng.upload(B.VEC_PREV, x0)
        prm = eng.newton_params(lin_reduction=1e-13)
        for _ in range(TRACER_STEPS):
            st, its, shift, a, s, u = eng.newton_step(prm)
            assert st == 0
            eng.advance_timestep()
        x = eng.download(B.VEC_CUR).reshape(hi - lo, nf)
        q.put((rank, {"x": x[b0 - lo:b1 - lo].copy(), "overlap_lo": x[0].copy(), "range": (lo, hi, b0, b1)}))
        eng.close()
    except BaseException:      # noqa: BLE001
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))
        raise


def test_slab_decomposed_tracer_matches_single_domain_oracle():
    import torch
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run under gpurun --gpus {world})")
    import torch.multiprocessing as mp
    from dumux_b200 import binding as B
    from oracle.oracle_py import Oracle
    ts = _tracer_problem()
    o = Oracle(ts)
    x = ts.initial.ravel().copy()
    for _ in range(TRACER_STEPS):
        r, j = o.assemble(x, x)
        dx, st, its, red = o.solve(j, r, reduction=1e-13)
        assert st == 0
        x = x - dx
    uid = B.Engine.nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tracer_worker, args=(r, world, uid, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    nf = TRACER_CELLS[0] * TRACER_CELLS[1]
    ref = x.reshape(TRACER_CELLS[2], nf)
    for r in range(world):
        assert "error" not in got[r], got[r].get("error")
        lo, hi, b0, b1 = got[r]["range"]
        assert np.abs(got[r]["x"] - ref[b0:b1]).max() <= 1e-12 * np.abs(ref).max()
        # the overlap layer carries the owner's values after every step (copyOwnerToAll inside the solve)
        assert np.abs(got[r]["overlap_lo"] - ref[lo]).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(ref - ts.initial.reshape(TRACER_CELLS[2], nf)).max() > 1e-13       # the field really moved
