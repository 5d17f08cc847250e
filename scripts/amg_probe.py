"""Per-level device time of the AMG V-cycle at 256^3 (or --cells): python scripts/amg_probe.py [--cells 256]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dumux_b200 import binding as B
from dumux_b200 import problems

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
c = args.cells
spec = problems.twop_lens((c, c, c), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
eng = B.Engine(spec, device=0)
eng.upload(B.VEC_PREV, spec.initial)
prm = eng.newton_params(lin_maxit=2000, preconditioner=B.PRECOND_AMG)
for it in range(2 + args.steps):
    if it == 2:
        eng.profile(True)
        eng.timer_start()
    eng.upload(B.VEC_CUR, spec.initial)
    st, its, shift, a, s, u = eng.newton_step(prm)
    assert st == 0
ms = eng.timer_stop() / args.steps
print(f"{c}^3 AMG-BiCGSTAB: {ms:.1f} ms per Newton step, {its} iterations, solve {s:.1f} ms")
tot = 0.0
for cells, (lms, n) in zip(eng.amg_levels(), eng.amg_level_profile()):
    print(f"  level {cells}: {lms:.3f} ms per cycle ({n} cycles)")
    tot += lms
print(f"  V-cycle total {tot:.3f} ms")
for name, k in (("spmv", B.K_SPMV), ("sweeps", B.K_ILU_APPLY), ("factor", B.K_ILU_FACTOR), ("blas1", B.K_BLAS1), ("amg_transfer", B.K_AMG)):
    t, n = eng.profile_read(k)
    print(f"  {name}: {t / args.steps:.2f} ms per step in {n // args.steps} units")
eng.close()
