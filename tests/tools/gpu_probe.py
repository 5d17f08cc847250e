"""Development probe (not a test): prints parity error magnitudes and kernel timings on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dumux_b200 import problems
from dumux_b200 import binding as B
from oracle.oracle_py import Oracle
import oracle.oracle_py as O


def parity(spec, name):
    o = Oracle(spec)
    rng = np.random.RandomState(3)
    cur = spec.initial.copy(); prev = spec.initial.copy()
    cur[:, 0] += rng.uniform(-50, 50, size=cur.shape[0])
    if spec.num_eq == 2:
        cur[:, 1] = rng.uniform(0, 0.3, size=cur.shape[0]); prev[:, 1] = rng.uniform(0, 0.3, size=cur.shape[0])
    ro, jo = o.assemble(cur, prev)
    e = B.Engine(spec)
    rg, jg = e.assemble(cur, prev)
    print(f"[{name}] res maxabs {np.abs(rg-ro).max():.3e} / {np.abs(ro).max():.3e}; jac maxabs {np.abs(jg-jo).max():.3e} / {np.abs(jo).max():.3e}; "
          f"bitexact res {np.array_equal(rg,ro)} jac {np.array_equal(jg,jo)} nbad {(jg!=jo).sum()}")
    xo, sto, ito, redo = o.solve(jo, ro)
    xg, stg, itg, redg = e.solve(jo, ro)
    print(f"[{name}] bicgstab oracle st {sto} it {ito} red {redo:.3e} | gpu st {stg} it {itg} red {redg:.3e} | dx rel {np.linalg.norm(xg-xo)/np.linalg.norm(xo):.3e}")
    e.close()


def timing(n):
    spec = problems.twop_lens((n, n, n), law="bc", heterogeneity_sigma=0.5)
    t = time.time(); e = B.Engine(spec); print(f"setup {n}^3: {time.time()-t:.1f}s")
    e.upload(B.VEC_CUR, spec.initial); e.upload(B.VEC_PREV, spec.initial)
    e.set_dt(250.0)
    e.assemble_device(True)
    cells = n ** 3
    for which, nm, bytes_per in ((B.KERNEL_ASSEMBLY, "assembly", 292), (B.KERNEL_SPMV, "spmv", 288)):
        ms = e.time_kernel(which, 5)
        print(f"{nm}: {ms:.3f} ms" + (f"  -> {bytes_per*cells/ms/1e6:.0f} GB/s algorithmic" if bytes_per else ""))
    t = time.time(); st = e.ilu0_factor(); e.synchronize(); print("ilu factor status", st, f"{time.time()-t:.3f}s")
    print("ilu factor ms", e.time_kernel(B.KERNEL_ILU_FACTOR, 2))
    print("ilu apply ms", e.time_kernel(B.KERNEL_ILU_APPLY, 3))
    p = e.newton_params()
    for i in range(3):
        st, its, shift, a, s, u = e.newton_step(p)
        print(f"newton step {i}: st {st} bicgstab its {its} shift {shift:.3e} assemble {a:.2f} ms solve {s:.2f} ms update {u:.2f} ms -> {cells*2/(max(a+s+u,1e-9)*1e-3)/1e6:.1f} MDOF/s")
    p = e.newton_params(preconditioner=B.PRECOND_BLOCKJACOBI, lin_maxit=2000)
    e.upload(B.VEC_CUR, spec.initial)
    for i in range(2):
        st, its, shift, a, s, u = e.newton_step(p)
        print(f"[jacobi] newton step {i}: st {st} bicgstab its {its} shift {shift:.3e} assemble {a:.2f} ms solve {s:.2f} ms update {u:.2f} ms -> {cells*2/(max(a+s+u,1e-9)*1e-3)/1e6:.1f} MDOF/s")
    e.close()


if __name__ == "__main__":
    parity(problems.onep_incompressible((20, 20)), "1p 2d")
    parity(problems.twop_lens((24, 16), law="vg"), "2p 2d vg")
    parity(problems.twop_lens((12, 10, 8), law="bc", heterogeneity_sigma=0.5), "2p 3d bc")
    for n in [int(a) for a in sys.argv[1:]] or [128]:
        timing(n)
