"""Development probe: ILU0 apply / factor / SpMV timings on the 2p lens problem.  usage: ilu_probe.py edge [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dumux_b200 import problems
from dumux_b200 import binding as B

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nz = int(sys.argv[3]) if len(sys.argv) > 3 else edge
spec = problems.twop_lens((edge, edge, nz), law="bc", heterogeneity_sigma=0.5, plane_rng=True)
e = B.Engine(spec)
e.upload(B.VEC_CUR, spec.initial); e.upload(B.VEC_PREV, spec.initial)
e.assemble_device(True)
e.ilu0_factor()
e.copy(B.VEC_WORK0, B.VEC_RESIDUAL)
n = edge * edge * nz
for which, nm, bpc in ((B.KERNEL_ILU_APPLY, "ilu apply", 288), (B.KERNEL_SPMV, "spmv", 288), (B.KERNEL_ILU_FACTOR, "ilu factor", 0), (B.KERNEL_ASSEMBLY, "assembly+volvars", 292)):
    e.time_kernel(which, 2)
    ms = e.time_kernel(which, reps)
    print(f"{edge}x{edge}x{nz} {nm}: {ms:.3f} ms" + (f" -> {bpc * n / ms / 1e6:.0f} GB/s" if bpc else ""))
e.close()
