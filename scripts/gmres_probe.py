"""Development probe: time per iteration of ILU0-GMRes(m) against ILU0-BiCGSTAB on the 2p lens problem.  usage: gmres_probe.py [edge=256] [its=40]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dumux_b200 import binding as B
from dumux_b200 import problems

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
its = int(sys.argv[2]) if len(sys.argv) > 2 else 40
spec = problems.twop_lens((edge, edge, edge), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
e = B.Engine(spec)
e.upload(B.VEC_CUR, spec.initial); e.upload(B.VEC_PREV, spec.initial)
import time
import numpy as np
e.assemble_device(True)
zero = np.zeros(e.n * e.b)
for kind, restart in (("bicgstab", 0), ("gmres", 10), ("gmres", 30)):
    e.set_linear_solver(kind, restart)
    e.upload(B.VEC_DELTA, zero)
    e.synchronize() if hasattr(e, "synchronize") else None
    t0 = time.perf_counter()
    st, n_it, red = e.solve_device(reduction=1e-30, maxit=its)
    s = (time.perf_counter() - t0) * 1e3          # host wall clock: the solve ends with a host read of the last scalars
    print(f"{edge}^3 {kind}({restart}): {n_it} iterations, solve {s:.1f} ms -> {s / max(n_it, 1):.2f} ms per iteration (operator applications: "
          f"{'2' if kind == 'bicgstab' else '1'} per iteration)")
e.close()
