"""Development probe: ILU0 apply / SpMV timings with 1x1 blocks (1p incompressible, log-normal K).  usage: ilu_probe_1p.py edge [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dumux_b200 import problems
from dumux_b200 import binding as B

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec = problems.onep_incompressible((edge, edge, edge), analytic=True)
spec.K = spec.K * problems.fast_lognormal_multiplier(spec.num_cells, 0.5, 0)
e = B.Engine(spec)
e.upload(B.VEC_CUR, spec.initial)
e.assemble_device(True)
e.ilu0_factor()
e.copy(B.VEC_WORK0, B.VEC_RESIDUAL)
n = edge ** 3
for which, nm, bpc in ((B.KERNEL_ILU_APPLY, "ilu apply", 104), (B.KERNEL_SPMV, "spmv", 104)):
    e.time_kernel(which, 2)
    ms = e.time_kernel(which, reps)
    print(f"1p {edge}^3 {nm}: {ms:.3f} ms -> {bpc * n / ms / 1e6:.0f} GB/s")
e.close()
