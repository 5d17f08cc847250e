"""Developer diagnostic: prints the per-chunk clock64 timeline of the structured ILU sweeps (DMX_SK_TRACE=1)."""
import os, sys
os.environ["DMX_SK_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dumux_b200 import problems
from dumux_b200 import binding as B
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SKC = int(sys.argv[2]) if len(sys.argv) > 2 else 8       # SK_CHUNK of the build
spec = problems.twop_lens((edge, edge, edge), law="bc", heterogeneity_sigma=0.5, plane_rng=True)
e = B.Engine(spec)
e.upload(B.VEC_CUR, spec.initial); e.upload(B.VEC_PREV, spec.initial)
e.assemble_device(True); e.ilu0_factor(); e.copy(B.VEC_WORK0, B.VEC_RESIDUAL)
for _ in range(3):
    e.ilu0_apply(B.VEC_WORK0, B.VEC_WORK1)
e.synchronize()
tr = e.sweep_trace()
names = ["lower", "upper"]
for kern in range(2):
    for tile in range(2):
        t = tr[kern, tile]
        nch = int((t[:, 0] != 0).sum())
        if nch == 0:
            continue
        t = t[:nch].astype(np.float64)
        t0 = t[0, 0]
        print(f"--- {names[kern]} sweep, tile slot {tile} ({'first' if tile == 0 else 'middle'} ticket), {nch} chunks, total {(t[-2, 19] - t0) / 1.9e3:.1f} us (at 1.9 GHz)")
        poll = t[:, 1] - t[:, 0]; halo = t[:, 2] - t[:, 1]
        wait = sum(t[:, 3 + 2 * c] - (t[:, 2] if c == 0 else t[:, 4 + 2 * (c - 1)]) for c in range(SKC))
        comp = sum(t[:, 4 + 2 * c] - t[:, 3 + 2 * c] for c in range(SKC))
        pub = t[:, 19] - t[:, 4 + 2 * (SKC - 1)]
        t = t[:-1]; poll = poll[:-1]; halo = halo[:-1]; wait = wait[:-1]; comp = comp[:-1]; pub = pub[:-1]
        for nm, arr in (("wait ready", poll), ("(unused)", halo), ("mbar_wait/chunk", wait), ("compute+barrier/chunk", comp), ("chunk tail", pub)):
            print(f"   {nm:22s} mean {arr.mean():9.0f} cyc  median {np.median(arr):9.0f}  max {arr.max():9.0f}")
        print("   chunk total mean", (t[:, 19] - t[:, 0]).mean(), "cycles")
        print("   first chunks:", [(int(t[c, 1] - t[c, 0]), int(t[c, 2] - t[c, 1]), int(t[c, 4 + 2 * (SKC - 1)] - t[c, 2]), int(t[c, 19] - t[c, 18])) for c in range(min(6, len(t)))])
e.close()
