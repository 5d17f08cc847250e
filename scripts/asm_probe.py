"""Development probe: device time of the assembly kernel (and the other graded kernels) at edge^3, 2p lens problem.
usage: python scripts/asm_probe.py [edge=256] [reps=10]   (DMX_ASM_SPLIT=0: one thread per cell)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dumux_b200 import binding as B
from dumux_b200 import problems

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec = problems.twop_lens((edge, edge, edge), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True, analytic=os.environ.get("ANALYTIC", "0") == "1")
e = B.Engine(spec)
n = spec.num_cells
rng = np.random.RandomState(5)
cur = spec.initial.copy()
cur[:, 1] = rng.uniform(0.0, 0.3, size=n)
e.upload(B.VEC_CUR, cur)
e.upload(B.VEC_PREV, spec.initial)
e.assemble_device(True)
ms = e.time_kernel(B.KERNEL_ASSEMBLY, reps)
bytes_ = n * (2 * 16 + 8 + 8 + 4 + 16) + e.nnzb * 32
print(f"assembly {edge}^3 split={os.environ.get('DMX_ASM_SPLIT', '1')}: {ms:.3f} ms, {bytes_ / ms / 1e6:.0f} GB/s algorithmic, frac {bytes_ / ms / 1e6 / 6538.9:.3f}")
res = e.download(B.VEC_RESIDUAL)
print("residual checksum", float(np.abs(res).sum()), "jac checksum", float(np.abs(e.jacobian()).sum()))
