"""one measured line per BASELINE configuration (C1, C2, C3(i)/(ii), C5), B200 arm next to the CPU oracle.

  python scripts/baseline_table.py [--edge 256] [--tracer-edge 512] [--out profiles/r5_baseline_table.jsonl]
  python -m torch.distributed.run --nproc-per-node N ... scripts/baseline_table.py --only c5 --tracer-edge 512      (C5 on N GPUs)

C3(i) at 256^3 and C4 (512^3, 1/2/4/8 GPUs) are bench.py's own lines (`value`, `strong_512`); this script adds the rest:
  C1  1p incompressible 2-D 100 x 100, stationary: assembly + ILU0-BiCGSTAB solve                       
  C2  1p compressible 3-D edge^3, tabulated water, log-normal K: the Newton iterations of the first time step
  C3ii  2p lens 3-D edge^3: the Newton iterations of the THIRD time step (plume present)
  C5  tracer 3-D: stationary pressure solve -> volume fluxes -> explicit and implicit transport steps; with torchrun the
      block-decomposed run (strong scaling)
Device times are CUDA-event times of dmx_newton_step's buckets; the CPU numbers time the oracle port on this box's host cores on
a bounded sample that is named in the line ("cpu": {...}).  Every rank > 0 only takes part in the collectives."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from dumux_b200 import binding as B
from dumux_b200 import problems

ap = argparse.ArgumentParser()
ap.add_argument("--edge", type=int, default=256)
ap.add_argument("--tracer-edge", type=int, default=512)
ap.add_argument("--cpu-edge", type=int, default=96)
ap.add_argument("--only", default="", help="comma-separated subset of c1,c2,c3,c5")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--out", default="")
args = ap.parse_args()
only = set(x for x in args.only.split(",") if x)

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))


def log(*a):
    if rank == 0:
        print(*a, file=sys.stderr, flush=True)


def nccl_uid():
    if dist is None:
        return None
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(B.Engine.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def maxreduce(x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


lines = []


def emit(obj):
    if rank == 0:
        lines.append(obj)
        print(json.dumps(obj), flush=True)


def want(c):
    return not only or c in only


def newton_iterations(eng, prm, max_steps=18, min_steps=2, allow_failure=False):
    """NewtonSolver::solveImpl_'s loop (newtonsolver.hh:976-1072) on the device state, one timed dmx_newton_step per iteration"""
    out = []
    steps, shift, last, conv = 0, 0.0, 0.0, False
    while True:
        if steps >= min_steps and (conv or (steps >= max_steps and not (shift * 4.0 < last))):
            break
        last = shift
        st, its, shift, a, s, u = eng.newton_step(prm)
        if st != 0:
            if allow_failure:
                return out, st
            raise SystemExit(f"Newton step failed: status {st}")
        out.append({"linear_iterations": its, "assemble_ms": a, "solve_ms": s, "update_ms": u, "shift": shift})
        steps += 1
        conv = shift <= 1e-8
    if allow_failure:
        return out, (0 if conv else 1)
    return out


def summary(its, dofs):
    tot = sum(i["assemble_ms"] + i["solve_ms"] + i["update_ms"] for i in its)
    return {"newton_iterations": len(its), "linear_iterations": [i["linear_iterations"] for i in its], "ms_total": tot,
            "ms_per_newton_iteration": tot / len(its), "mdof_per_s": dofs * len(its) / (tot * 1e-3) / 1e6,
            "assemble_ms": [round(i["assemble_ms"], 3) for i in its], "solve_ms": [round(i["solve_ms"], 2) for i in its]}


# ---------------------------------------------------------------------------------------------------------------------------
if want("c1") and world == 1:
    spec = problems.onep_incompressible((100, 100))
    e = B.Engine(spec)
    prm = e.newton_params()
    runs = []
    for i in range(6):
        e.upload(B.VEC_CUR, spec.initial)
        st, its, shift, a, s, u = e.newton_step(prm)
        assert st == 0
        if i:
            runs.append((a, s, u, its))
    a, s, u = (float(np.mean([r[k] for r in runs])) for k in range(3))
    gpu = {"linear_solver": "ILU0-BiCGSTAB", "linear_iterations": runs[-1][3], "assemble_ms": a, "solve_ms": s, "update_ms": u,
           "mdof_per_s": 1e4 / ((a + s + u) * 1e-3) / 1e6}
    e.close()
    cpu = None
    if not args.no_cpu:
        from oracle.oracle_py import Oracle
        o = Oracle(spec, num_threads=1)
        u0 = spec.initial.reshape(-1).copy()
        t0 = time.perf_counter()
        for _ in range(5):
            res, jac = o.assemble(u0, u0)
            dx, st, its, red = o.solve(jac, res, reduction=1e-6, maxit=250)
        sec = (time.perf_counter() - t0) / 5
        cpu = {"what": "oracle port, the full configuration, 1 core", "linear_iterations": its, "ms": sec * 1e3, "mdof_per_s": 1e4 / sec / 1e6}
    emit({"config": "C1", "workload": "1p incompressible CCTpfa 2-D 100x100, stationary: assembly + ILU0-BiCGSTAB(1e-6) + update", "gpus": 1,
          "gpu": gpu, "cpu": cpu})

# ---------------------------------------------------------------------------------------------------------------------------
if want("c2") and world == 1:
    E = args.edge
    cells = (E, E, E)
    t0 = time.time()
    spec = problems.onep_compressible(cells, lognormal=True, dt=0.002)
    log(f"[C2] host set-up {time.time() - t0:.1f} s")
    gpu = {}
    for solver in ("ilu0", "amg"):
        e = B.Engine(spec)
        prm = e.newton_params(lin_maxit=2000, preconditioner=B.PRECOND_AMG if solver == "amg" else B.PRECOND_ILU0)
        for rep in range(2):                      # the first pass pays one-time set-up (table upload, AMG hierarchy)
            e.upload(B.VEC_CUR, spec.initial)
            e.upload(B.VEC_PREV, spec.initial)
            its = newton_iterations(e, prm)
        gpu[solver] = summary(its, E ** 3)
        gpu[solver]["assembly_gbs_of_100B_per_cell"] = 100 * E ** 3 / (its[-1]["assemble_ms"] * 1e-3) / 1e9
        e.close()
    cpu = None
    if not args.no_cpu:
        from oracle.oracle_py import Oracle
        ce = args.cpu_edge
        cs = problems.onep_compressible((ce, ce, ce), lognormal=True, dt=0.002)
        o = Oracle(cs, num_threads=os.cpu_count() or 1)
        u0 = cs.initial.reshape(-1).copy()
        t0 = time.perf_counter()
        res, jac = o.assemble(u0, u0)
        dx, st, its_c, red = o.solve(jac, res, reduction=1e-6, maxit=2000)
        sec = time.perf_counter() - t0
        cpu = {"what": f"oracle port, bounded sample {ce}^3, first Newton iteration, assembly on {os.cpu_count()} threads, ILU0-BiCGSTAB on 1 core",
               "linear_iterations": its_c, "ms": sec * 1e3, "mdof_per_s": ce ** 3 / sec / 1e6}
    emit({"config": "C2", "workload": f"1p compressible CCTpfa 3-D {E}^3, TabulatedComponent<H2O>, log-normal K, dt 0.002 s: the Newton iterations of the first time step",
          "gpus": 1, "gpu": gpu, "cpu": cpu})
    del spec

# ---------------------------------------------------------------------------------------------------------------------------
if want("c3") and world == 1:
    E = args.edge
    cells = (E, E, E)
    spec = problems.twop_lens(cells, law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
    gpu = {}
    for solver in ("ilu0", "amg"):
        e = B.Engine(spec)
        prm = e.newton_params(lin_maxit=2000, preconditioner=B.PRECOND_AMG if solver == "amg" else B.PRECOND_ILU0)
        e.upload(B.VEC_CUR, spec.initial)
        e.upload(B.VEC_PREV, spec.initial)
        per_step = []
        dt = 250.0
        for tstep in range(3):
            # NewtonSolver::solve (newtonsolver.hh:309-355): a failed Newton solve (linear solver breakdown, no convergence)
            # resets the solution and halves dt; dt stays at 250 s otherwise
            failures = []
            for attempt in range(6):
                e.set_dt(dt)
                its, st = newton_iterations(e, prm, allow_failure=True)
                if st == 0:
                    break
                failures.append({"dt": dt, "status": st, "newton_iterations_before_failure": len(its),
                                 "linear_iterations": [i["linear_iterations"] for i in its]})
                e.reset_timestep()
                dt *= 0.5
            else:
                raise SystemExit(f"C3 {solver}: time step {tstep + 1} failed down to dt {dt}")
            sm = summary(its, 2 * E ** 3)
            sm["dt"] = dt
            sm["failed_attempts"] = failures
            per_step.append(sm)
            e.advance_timestep()
        sn = e.download(B.VEC_CUR).reshape(-1, 2)[:, 1]
        gpu[solver] = {"time_step_1": per_step[0], "time_step_3": per_step[2], "napl_cells_after_3_steps": int((sn > 1e-6).sum()),
                       "max_S_n": float(sn.max())}
        e.close()
    emit({"config": "C3", "workload": f"2p immiscible CCTpfa lens {E}^3 (bench.py's workload): (i) the Newton iterations of the first time step, "
                                      f"(ii) of the third time step (plume present), dt 250 s", "gpus": 1, "gpu": gpu,
          "cpu": "bench.py --impl reference (the full 256^3 configuration, first Newton iteration)"})
    del spec

# ---------------------------------------------------------------------------------------------------------------------------
if want("c5"):
    E = args.tracer_edge
    cells = (E, E, E)
    part = {1: None, 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world, None)
    if world > 1:
        rng = problems.box_partition(cells, part, rank)
        box = [(r[0], r[1]) for r in rng]
    else:
        box = None
    t0 = time.time()
    ps = problems.onep_tracer_pressure_large(cells, box=box)
    log(f"[C5] host set-up of the pressure problem {time.time() - t0:.1f} s")
    e1 = B.Engine(ps, device=local_rank, nccl_uid=nccl_uid(), rank=rank, nranks=world, part=part)
    e1.set_linear_solver("cg")
    prm = e1.newton_params(lin_maxit=2000, lin_reduction=1e-10, preconditioner=B.PRECOND_AMG)
    e1.upload(B.VEC_CUR, ps.initial)
    st, its, shift, a, s, u = e1.newton_step(prm)
    assert st == 0, st
    pressure = {"linear_solver": "AMG-CG (AMGCGIstlSolver, examples/1ptracer/main.cc:125-133), reduction 1e-10", "linear_iterations": its,
                "assemble_ms": a, "solve_ms": maxreduce(s)}
    p = e1.download(B.VEC_CUR)
    vf = e1.volume_flux(p)
    e1.close()
    del ps, p
    n_global = E ** 3
    tr = {}
    for implicit in (False, True):
        ts = problems.tracer_transport(cells, vf, dt=10.0 if not implicit else 100.0, implicit=implicit, box=box)
        et = B.Engine(ts, device=local_rank, nccl_uid=nccl_uid(), rank=rank, nranks=world, part=part)
        et.upload(B.VEC_CUR, ts.initial)
        et.upload(B.VEC_PREV, ts.initial)
        prm = et.newton_params(lin_reduction=1e-10, lin_maxit=500)
        steps = 6
        et.synchronize()
        for i in range(steps):
            if i == 1:
                et.synchronize()
                if dist is not None:
                    dist.barrier()
                et.timer_start()
                buckets, nit = [0.0, 0.0, 0.0], 0
            st, its, shift, a, s, u = et.newton_step(prm)
            assert st == 0
            et.advance_timestep()
            if i:
                buckets[0] += a; buckets[1] += s; buckets[2] += u; nit += its
        ms = maxreduce(et.timer_stop()) / (steps - 1)
        x = et.download(B.VEC_CUR)
        tr["implicit" if implicit else "explicit"] = {
            "dt": 10.0 if not implicit else 100.0, "ms_per_time_step": ms, "mdof_per_s": n_global / (ms * 1e-3) / 1e6,
            "assemble_ms": buckets[0] / (steps - 1), "solve_ms": buckets[1] / (steps - 1), "update_ms": buckets[2] / (steps - 1),
            "linear_iterations_per_step": nit / (steps - 1), "max_mass_fraction_rank0": float(x.max()),
            "assembly_gbs_of_136B_per_cell_rank0": 136 * et.n / (buckets[0] / (steps - 1) * 1e-3) / 1e9}
        et.close()
        del ts
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle.oracle_py import Oracle
        ce = args.cpu_edge
        cps = problems.onep_tracer_pressure_large((ce, ce, ce))
        o1 = Oracle(cps, num_threads=os.cpu_count() or 1)
        rng_ = np.random.RandomState(0)
        ctr = problems.cell_centers((ce, ce, ce), cps.lower, cps.upper)
        pp = 1.0e5 * (1.1 - 0.1 * ctr[:, 2]) + rng_.uniform(-20.0, 20.0, size=ce ** 3)
        cvf = o1.volume_flux(pp)
        cts = problems.tracer_transport((ce, ce, ce), cvf, dt=0.01)
        o = Oracle(cts, num_threads=os.cpu_count() or 1)
        x0 = cts.initial.reshape(-1).copy()
        t0 = time.perf_counter()
        nst = 3
        for _ in range(nst):
            r, j = o.assemble(x0, x0)
            dx, st, its_c, red = o.solve(j, r, reduction=1e-10)
            x0 = x0 - dx
        sec = (time.perf_counter() - t0) / nst
        cpu = {"what": f"oracle port, bounded sample {ce}^3, explicit tracer step (assembly on {os.cpu_count()} threads + diagonal solve on 1 core)",
               "ms": sec * 1e3, "mdof_per_s": ce ** 3 / sec / 1e6}
    emit({"config": "C5", "workload": f"tracer transport CCTpfa 3-D {E}^3 on the frozen velocity field of the stationary 1p solve (examples/1ptracer at benchmark size)",
          "gpus": world, "partitioning": list(part) if part else [1, 1, 1], "pressure_solve": pressure, "tracer": tr, "cpu": cpu})

if rank == 0 and args.out:
    with open(os.path.join(ROOT, args.out), "a") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
