#!/usr/bin/env python
"""bench.py -- Newton-step throughput of the B200 CCTpfa engine (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm   (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (the oracle port) on host cores

One "step" = one Newton iteration of the 2p immiscible CCTpfa lens problem: residual + Jacobian
assembly by numeric differentiation, fresh ILU0 factorisation, BiCGSTAB to LinearSolver.ResidualReduction = 1e-6,
Newton update + shift -- the loop body of NewtonSolver::solveImpl_ (dumux/nonlinear/newtonsolver.hh:998-1062).
Every step starts from the same state (hydrostatic initial condition, first iteration of the first time step), so
all K steps do identical work.  metric = degrees of freedom (cells x 2) / second, in MDOF/s.

N = 1: 256^3 cells (the configuration the metric is quoted on).  N > 1: weak scaling, 256 x 256 x (256 N) cells slab-
decomposed along z with overlap 1 (Grid.Partitioning "1 1 N"), per-rank ILU0 (overlapping Schwarz) -- pass --cells to
change the per-GPU cube edge, --global-z to fix the total number of layers instead (strong scaling).

The JSON line also carries: `roofline` (dominant roofline-graded kernel, CUDA-event timed inside the timed region),
`kernels` (every kernel class the same way), `e2e` (the same step through the host-buffer C-ABI call: pinned host
curSol -> H2D -> step -> D2H), `cpu_baseline` (the oracle port on this box's host cores, bounded sample), `clocks`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Newton-step MDOF/s (assembly+BiCGSTAB) 2p CCTpfa"
UNIT = "MDOF/s"
LIN_MAXIT = 2000          # LinearSolver.MaxIterations: ILU0-BiCGSTAB needs > 250 iterations at 256^3 (no AMG on this path)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write banners to file descriptor 1 (NCCL prints "NCCL version ..." at
# communicator creation when NCCL_DEBUG=VERSION is set in the environment), so fd 1 is pointed at stderr for the whole run and
# the result line goes to a duplicate of the original stdout.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


# ----------------------------------------------------------------------------------------------------------
# algorithmic bytes per launch (DESIGN.md "Kernels")
# ----------------------------------------------------------------------------------------------------------
def algorithmic_bytes(n, nnzb, b):
    vec = n * b * 8
    return {
        # read cur, prev, K, phi, region; write residual + Jacobian blocks (neighbour reads are cache hits)
        "assembly": n * (2 * b * 8 + 8 + 8 + 4 + b * 8) + nnzb * b * b * 8,
        # structured-grid SpMV (no column indices): values per block, rowptr + x + y per row -- what the kernel must move
        "spmv": nnzb * 8 * b * b + n * (4 + 16 * b),
        # the CSR/BCRS-equivalent figure of SURVEY 8d (values + colidx per block, rowptr + x + y per row), reported alongside
        "spmv_csr_equivalent": nnzb * (8 * b * b + 4) + n * (4 + 16 * b),
        # factors + colidx, d read, v written / re-read / written
        "ilu0_apply": nnzb * (8 * b * b + 4) + n * (4 + 4 * 8 * b),
        "vec": vec,
    }


# ----------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx), "-f", self.path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


# ----------------------------------------------------------------------------------------------------------
# CPU side: the oracle port (reference algorithm restated), bounded sample of the same workload
# ----------------------------------------------------------------------------------------------------------
def cpu_newton_step(edge, steps, warmup, threads):
    """Times `steps` Newton iterations (assemble + ILU0 + BiCGSTAB + update) of the 2p lens problem at edge^3 cells with
    the oracle.  Assembly runs on `threads` OpenMP threads (DuMux's coloured parallelFor), the linear solve is the
    sequential dune-istl algorithm (one thread per rank, as in the reference)."""
    import numpy as np
    from dumux_b200 import problems
    from oracle.oracle_py import Oracle
    spec = problems.twop_lens((edge, edge, edge), law="bc", heterogeneity_sigma=0.5, plane_rng=True)
    o = Oracle(spec, num_threads=threads)
    u0 = spec.initial.reshape(-1).copy()
    times, its_seen = [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res, jac = o.assemble(u0, u0)
        dx, st, its, red = o.solve(jac, res, reduction=1e-6, maxit=LIN_MAXIT)
        u = u0 - dx
        shift = float(np.max(np.abs(u - u0) / np.maximum(1.0, np.abs(u + u0) * 0.5)))
        dt = time.perf_counter() - t0
        assert st == 0 and shift > 0
        if i >= warmup:
            times.append(dt)
            its_seen.append(its)
    dofs = 2 * edge ** 3
    sec = sum(times) / len(times)
    return {"value": dofs / sec / 1e6, "sec_per_step": sec, "bicgstab_iterations": its_seen[-1], "dofs": dofs}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    edge = args.cpu_edge
    r = cpu_newton_step(edge, args.steps, args.warmup, cores)
    sample = (f"2p lens {edge}^3 cells (bounded sample of the {args.cells}^3 workload), one Newton iteration per step, "
              f"{r['bicgstab_iterations']} BiCGSTAB iterations; oracle port of the DuMux/dune-istl algorithm: assembly on {cores} "
              f"OpenMP threads, ILU0/BiCGSTAB sequential (1 rank, no MPI in this image)")
    line = {
        "impl": "reference", "metric": f"{METRIC} {args.cells}^3", "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the arm's workload (same wording as the B200 arm's `config`), timed on a bounded sample of it
        "config": {"workload": f"2p immiscible CCTpfa lens/infiltration, {args.cells}x{args.cells}x{args.cells} cells ({args.cells}^3 per GPU), "
                               f"Brooks-Corey, lognormal K multiplier sigma 0.5, numeric differentiation (forward, eps 1e-10), 2x2 BCRS blocks",
                   "step": "one Newton iteration: assemble + ILU0 factor + BiCGSTAB(1e-6) + update, from the hydrostatic initial state, dt 250 s",
                   "linear_solver": f"ILU0-BiCGSTAB, reduction 1e-6, maxit {LIN_MAXIT}", "bicgstab_iterations_per_step": r["bicgstab_iterations"],
                   "parallelism": "host cores of the GPU box", "sample_cells": edge ** 3},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    from dumux_b200 import problems
    from dumux_b200 import binding as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(B.Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    edge = args.cells
    nz_global = args.global_z if args.global_z else edge * world
    cells = (edge, edge, nz_global)
    # physical domain: the C3 box [0,6]x[0,4]x[0,4] per 256-layer cube, stretched in z with the number of layers
    upper = (6.0, 4.0, 4.0 * nz_global / edge)
    lo, hi, b0, b1 = problems.slab_partition(nz_global, world, rank)
    t0 = time.time()
    spec = problems.twop_lens(cells, law="bc", upper=upper, lower=(0.0, 0.0, 0.0), heterogeneity_sigma=0.5, dt=250.0,
                              slab=(lo, hi) if world > 1 else None, plane_rng=True)
    eng = B.Engine(spec, device=local_rank, nccl_uid=uid, rank=rank, nranks=world)
    n_local, b = eng.n, eng.b
    n_owned = edge * edge * (b1 - b0)
    dofs_global = edge * edge * nz_global * b
    if rank == 0:
        log(f"[bench] set-up {time.time() - t0:.1f} s: global cells {cells}, rank 0 holds {n_local} cells ({n_owned} owned), nnzb {eng.nnzb}")

    u0 = torch.from_numpy(np.ascontiguousarray(spec.initial.reshape(-1))).pin_memory()
    uh = torch.empty_like(u0).pin_memory()
    eng.upload(B.VEC_PREV, u0)
    eng.upload(B.VEC_WORK1, u0)          # device copy of the start state
    prm = eng.newton_params(lin_maxit=LIN_MAXIT)

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def maxreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_step():
        eng.copy(B.VEC_CUR, B.VEC_WORK1)
        st, its, shift, a, s, u = eng.newton_step(prm)
        if st != 0:
            raise SystemExit(f"bench.py: Newton step failed with status {st}")
        return its, shift, a, s, u

    def host_step():
        uh.copy_(u0)                        # the caller's curSol lives on the host
        st, its, shift = eng.newton_step_host(uh, prm)
        if st != 0:
            raise SystemExit(f"bench.py: host Newton step failed with status {st}")
        return its, shift

    # ---- warm-up ----
    for _ in range(args.warmup):
        its, shift, *_ = device_step()
    if rank == 0:
        log(f"[bench] warm-up done: {its} BiCGSTAB iterations per step, shift {shift:.3e}")

    # ---- timed region: K device-resident steps ----
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = eng.launches()
    eng.profile(True)
    sampler.start()
    barrier()
    eng.timer_start()
    wall0 = time.perf_counter()
    buckets = [0.0, 0.0, 0.0]
    for _ in range(args.steps):
        its, shift, a, s, u = device_step()
        buckets[0] += a; buckets[1] += s; buckets[2] += u
    ms_total = eng.timer_stop()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = eng.launches() - launches0
    prof = {k: eng.profile_read(v) for k, v in (("assembly", B.K_ASSEMBLY), ("spmv", B.K_SPMV),
                                                ("ilu0_apply", B.K_ILU_APPLY), ("ilu0_factor", B.K_ILU_FACTOR),
                                                ("blas1", B.K_BLAS1), ("halo", B.K_HALO))}
    eng.profile(False)
    ms_total = maxreduce(ms_total)
    ms_per_step = ms_total / args.steps
    value = dofs_global / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the same step through the host-buffer ABI call ----
    for _ in range(2):
        host_step()
    barrier()
    eng.timer_start()
    e2e_steps = max(2, min(args.steps, 3))
    for _ in range(e2e_steps):
        host_step()
    ms_e2e = maxreduce(eng.timer_stop()) / e2e_steps
    barrier()
    e2e_value = dofs_global / (ms_e2e * 1e-3) / 1e6
    assert float((uh - u0).abs().max()) > 0.0           # the step really came back to the host

    # ---- roofline of the graded kernels, from the in-region event timers ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ab = algorithmic_bytes(n_local, eng.nnzb, b)
    # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/traffic.json), which was taken on the default
    # 256^3 single-GPU workload: reported only when this run has that per-GPU size, null otherwise
    traffic = {}
    if edge == 256 and n_local == 256 ** 3:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
    kernels = {}
    for name, (ms, units) in prof.items():
        if units == 0:
            continue
        avg = ms / units
        k = {"launches_timed": units, "avg_ms": avg, "share_of_step": ms / ms_total if ms_total > 0 else None}
        if name in ab:
            ach = ab[name] / (avg * 1e-3) / 1e9
            k.update(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, algorithmic_bytes=ab[name],
                     traffic=traffic.get(name))
            if name == "spmv":
                csr = ab["spmv_csr_equivalent"]
                k.update(csr_equivalent_bytes=csr, csr_equivalent_gbs=csr / (avg * 1e-3) / 1e9,
                         note="read-dominated stream: the peak is the measured COPY bandwidth (half writes), pure reads run above it; "
                              "against the nominal 8000 GB/s the fraction is %.2f" % (ach / 8000.0))
        kernels[name] = k
    # the dominant kernel of the step (largest share of the timed region) among the HBM-bound kernel classes
    graded = [k for k in ("ilu0_apply", "spmv", "assembly") if k in kernels]
    dom = max(graded, key=lambda k: kernels[k]["share_of_step"]) if graded else None
    roofline = None
    if dom:
        kd = kernels[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kd["achieved"], "peak": peak, "unit": "GB/s", "frac": kd["frac"],
                    "traffic": kd.get("traffic"), "peak_source": peak_src, "algorithmic_bytes": kd["algorithmic_bytes"],
                    "avg_launch_ms": kd["avg_ms"], "share_of_step": kd["share_of_step"],
                    "note": ("one ILU0 application = vec_skew + lower sweep + upper sweep (3 launches timed as one unit); "
                             "algorithmic bytes are the BCRS-equivalent figure of SURVEY 8d" if dom == "ilu0_apply" else None)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        eng.close()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t0 = time.time()
        r = cpu_newton_step(args.cpu_edge, 1, 0, cores)
        cpu = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"one Newton iteration of the same 2p lens problem at {args.cpu_edge}^3 cells ({r['bicgstab_iterations']} BiCGSTAB "
                         f"iterations, {r['sec_per_step']:.1f} s): oracle port, assembly on {cores} OpenMP threads, ILU0/BiCGSTAB sequential"}
        log(f"[bench] cpu baseline {time.time() - t0:.1f} s")

    line = {
        "metric": f"{METRIC} {edge}^3", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.global_z else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"2p immiscible CCTpfa lens/infiltration, {cells[0]}x{cells[1]}x{cells[2]} cells ({edge}^3 per GPU), Brooks-Corey, "
                               f"lognormal K multiplier sigma 0.5, numeric differentiation (forward, eps 1e-10), 2x2 BCRS blocks",
                   "step": "one Newton iteration: assemble + ILU0 factor + BiCGSTAB(1e-6) + update, from the hydrostatic initial state, dt 250 s",
                   "linear_solver": f"ILU0-BiCGSTAB, reduction 1e-6, maxit {LIN_MAXIT}", "bicgstab_iterations_per_step": its,
                   "parallelism": f"slab z x{world}, overlap 1" if world > 1 else "single GPU",
                   "l2_policy": "inputs larger than L2 (Jacobian 3.75 GB, vectors 268 MB per GPU at 256^3)"},
        "buckets_ms_per_step": {"assemble": buckets[0] / args.steps, "solve": buckets[1] / args.steps, "update": buckets[2] / args.steps},
        "wall_ms_per_step": wall * 1e3 / args.steps,
        "roofline": roofline, "kernels": kernels,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(u0.numel() * 8), "d2h_bytes_per_step": int(u0.numel() * 8),
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=256, help="cube edge per GPU")
    ap.add_argument("--global-z", type=int, default=0, help="fix the global number of z layers (strong scaling)")
    ap.add_argument("--cpu-edge", type=int, default=96, help="cube edge of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
