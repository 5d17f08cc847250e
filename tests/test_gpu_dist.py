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


def _make_spec(slab):
    from dumux_b200 import problems
    return problems.twop_lens(CELLS, law="bc", heterogeneity_sigma=0.4, slab=slab, plane_rng=True)


def _perturb(spec, slab, seed=3):
    n = int(np.prod(CELLS))
    rng = np.random.RandomState(seed)
    sn = rng.uniform(0.0, 0.25, size=n)
    dp = rng.uniform(-40.0, 40.0, size=n)
    if slab is not None:
        nf = CELLS[0] * CELLS[1]
        sn = sn.reshape(CELLS[2], nf)[slab[0]:slab[1]].reshape(-1)
        dp = dp.reshape(CELLS[2], nf)[slab[0]:slab[1]].reshape(-1)
    u = spec.initial.copy()
    u[:, 0] += dp
    u[:, 1] = sn
    return u


def _cpu_rank_job(rank_obj):
    slab = (rank_obj.lo, rank_obj.hi) if rank_obj.comm.nranks > 1 else None
    cur = _perturb(rank_obj.spec, slab).reshape(-1)
    prev = rank_obj.spec.initial.reshape(-1)
    res, jac = rank_obj.o.assemble(cur, prev)
    x, st, its, red = rank_obj.bicgstab(jac, res, reduction=1e-10, maxit=500)
    u, nst, nsteps, lin_its = rank_obj.newton(rank_obj.spec.initial, rank_obj.spec.initial)
    return {"res": res, "jac": jac, "x": x, "st": st, "its": its, "u": u, "nst": nst, "nsteps": nsteps, "lin_its": lin_its}


def _gpu_worker(rank, world, uid, q):
    try:
        from dumux_b200 import binding as B
        from dumux_b200 import problems
        lo, hi, b0, b1 = problems.slab_partition(CELLS[2], world, rank)
        spec = _make_spec((lo, hi))
        eng = B.Engine(spec, device=rank, nccl_uid=uid, rank=rank, nranks=world)
        assert (eng.own_begin, eng.own_end) == (b0 - lo, b1 - lo)
        cur = _perturb(spec, (lo, hi))
        res, jac = eng.assemble(cur, spec.initial)
        x, st, its, red = eng.solve(jac, res, reduction=1e-10, maxit=500)
        # halo exchange primitive: fill a vector with the rank id, exchange, look at the overlap planes
        v = np.full(eng.n * eng.b, float(rank))
        eng.upload(B.VEC_WORK1, v)
        eng.halo_exchange(B.VEC_WORK1)
        halo = eng.download(B.VEC_WORK1).reshape(hi - lo, -1)[:, 0].copy()
        nrm = eng.norm(B.VEC_RESIDUAL)
        u, nst, rep = eng.newton(spec.initial, spec.initial)
        q.put((rank, {"res": res, "jac": jac, "x": x, "st": st, "its": its, "halo": halo, "norm": nrm, "u": u, "nst": nst,
                      "nsteps": rep.newton_iterations, "lin_its": [rep.linear_iterations[i] for i in range(rep.newton_iterations)],
                      "launches": eng.launches()}))
        eng.close()
    except BaseException as e:      # noqa: BLE001
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))
        raise


@pytest.mark.parametrize("world", [2])
def test_slab_decomposed_newton_step_matches_cpu_reference(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run under gpurun --gpus {world})")
    import torch.multiprocessing as mp
    from dumux_b200 import binding as B
    from dumux_b200 import problems
    from oracle import dist_oracle as D
    ref = D.run_threads(_make_spec, CELLS, world, _cpu_rank_job)
    uid = B.Engine.nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, uid, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert "error" not in got[r], got[r].get("error")
    nf2 = CELLS[0] * CELLS[1] * 2
    for r in range(world):
        g, c = got[r], ref[r]
        lo, hi, b0, b1 = problems.slab_partition(CELLS[2], world, r)
        # assembly of the local box incl. overlap rows (no scvf on the processor boundary): bit-exact
        assert np.array_equal(g["res"], c["res"]) and np.array_equal(g["jac"], c["jac"])
        # halo: overlap planes carry the neighbour's rank id, owned planes mine
        expect = np.full(hi - lo, float(r))
        if lo > 0:
            expect[0] = r - 1
        if hi < CELLS[2]:
            expect[-1] = r + 1
        assert np.array_equal(g["halo"], expect)
        # owner-masked, all-reduced norm is the same number on every rank
        assert g["norm"] == got[0]["norm"]
        # Schwarz-BiCGSTAB: same iteration count, solution at the solver tolerance
        assert g["st"] == 0 and c["st"] == 0 and g["its"] == c["its"], (g["its"], c["its"])
        assert np.linalg.norm(g["x"] - c["x"]) <= 1e-7 * np.linalg.norm(c["x"])
        # Newton: same iteration count, fields to 1e-8
        assert g["nst"] == 0 and g["nsteps"] == c["nsteps"]
        ug, uc = g["u"].reshape(-1, 2), c["u"].reshape(-1, 2)
        assert np.linalg.norm(ug[:, 0] - uc[:, 0]) <= 1e-8 * np.linalg.norm(uc[:, 0])
        assert np.linalg.norm(ug[:, 1] - uc[:, 1]) <= 1e-8 * max(1.0, np.linalg.norm(uc[:, 1]))
        assert g["launches"] > 0
    # global norm = sqrt(sum of owned squares)
    owned = D.gather_owned([ref[r]["res"] for r in range(world)], CELLS, world, 2)
    assert abs(got[0]["norm"] - np.linalg.norm(owned)) <= 1e-12 * np.linalg.norm(owned)
