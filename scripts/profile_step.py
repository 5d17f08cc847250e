"""Short single-GPU driver for ncu captures (not a test, not a bench): one assembly, one ILU0 factorisation and a few
BiCGSTAB iterations of the 2p lens problem, so that every hot kernel is launched a handful of times.
usage: python scripts/profile_step.py [edge=256] [bicgstab_iterations=3] [amg]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dumux_b200 import binding as B
from dumux_b200 import problems

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
its = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = problems.twop_lens((edge, edge, edge), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
e = B.Engine(spec)
e.upload(B.VEC_CUR, spec.initial)
e.upload(B.VEC_PREV, spec.initial)
p = e.newton_params(lin_maxit=its)
if len(sys.argv) > 3 and sys.argv[3] == "amg":
    p.preconditioner = B.PRECOND_AMG
st, n_it, shift, a, s, u = e.newton_step(p)
print(f"status {st} (1 = stopped at maxit, expected), {n_it} BiCGSTAB iterations, {e.launches()} launches")
e.close()
