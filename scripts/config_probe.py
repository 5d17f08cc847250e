"""Development probe (not a test, not the bench): device timings of the other BASELINE configurations on one B200.
  config 2: 1p compressible (tabulated water), log-normal K, one Newton iteration
  config 5: tracer transport on the frozen velocity field of a stationary 1p solve: volume-flux kernel + explicit/implicit steps
usage: python scripts/config_probe.py [edge=256]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dumux_b200 import binding as B
from dumux_b200 import problems

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cells = (edge, edge, edge)
n = edge ** 3

t0 = time.time()
spec = problems.onep_compressible(cells, lognormal=True, dt=0.002)
print(f"[config 2] host set-up {time.time() - t0:.1f} s")
e = B.Engine(spec)
e.upload(B.VEC_CUR, spec.initial); e.upload(B.VEC_PREV, spec.initial)
prm = e.newton_params(lin_maxit=2000)
for i in range(3):
    st, its, shift, a, s, u = e.newton_step(prm)
    print(f"[config 2] 1p compressible {edge}^3 Newton iteration {i}: status {st}, {its} BiCGSTAB its, shift {shift:.2e}, assemble {a:.2f} ms "
          f"({100 * n / a / 1e6:.0f} GB/s of 100 B/cell), solve {s:.1f} ms, update {u:.2f} ms -> {n / ((a + s + u) * 1e-3) / 1e6:.1f} MDOF/s")
e.close()

t0 = time.time()
ps = problems.onep_tracer_pressure(cells) if n <= 4_000_000 else None
if ps is None:
    # large grids: same problem with the fast log-normal field (the mt19937 replay is a Python loop)
    ps = problems.onep_tracer_pressure((4, 4, 4))
    ps = problems.ProblemSpec(**{**ps.__dict__, "cells": cells})
    ctr = problems.cell_centers(cells, ps.lower, ps.upper)
    lens = problems._in_box(ctr, [0.2] * 3, [0.8] * 3, 1.5e-7)
    ps.K = np.where(lens, 1e-11, 1e-10) * problems.fast_lognormal_multiplier(n, 0.5, 0)
    ps.phi = np.full(n, 0.2); ps.region = np.zeros(n, dtype=np.int32); ps.initial = np.zeros((n, 1))
    zmax = 1.0
    for side in range(6):
        fc = problems.side_face_centers(cells, ps.lower, ps.upper, side)
        z = fc[:, 2]
        d = (z < 1e-6) | (z > zmax - 1e-6)
        ps.bc_type[side] = np.where(d, problems.BC_DIRICHLET, problems.BC_NEUMANN).astype(np.int32)
        v = np.zeros((fc.shape[0], 1)); v[d, 0] = 1.0e5 * (1.1 - z[d] * 0.1); ps.bc_values[side] = v
print(f"[config 5] host set-up of the 1p problem {time.time() - t0:.1f} s")
e1 = B.Engine(ps)
e1.upload(B.VEC_CUR, ps.initial)
prm = e1.newton_params(lin_maxit=4000, lin_reduction=1e-10)
st, its, shift, a, s, u = e1.newton_step(prm)
print(f"[config 5] stationary 1p pressure solve {edge}^3: status {st}, {its} BiCGSTAB its, assemble {a:.2f} ms, solve {s:.1f} ms")
p = e1.download(B.VEC_CUR)
t0 = time.time(); vf = e1.volume_flux(p); print(f"[config 5] volume fluxes incl. D2H of {vf.nbytes / 1e9:.2f} GB: {time.time() - t0:.2f} s")
e1.close()
for implicit in (False, True):
    ts = problems.tracer_transport(cells, vf, dt=10.0 if not implicit else 100.0, implicit=implicit)
    et = B.Engine(ts)
    et.upload(B.VEC_CUR, ts.initial); et.upload(B.VEC_PREV, ts.initial)
    prm = et.newton_params(lin_reduction=1e-10, lin_maxit=500)
    tot = [0.0, 0.0, 0.0]; nit = 0
    steps = 5
    for i in range(steps):
        st, its, shift, a, s, u = et.newton_step(prm)
        et.advance_timestep()
        if i:
            tot[0] += a; tot[1] += s; tot[2] += u; nit += its
    k = steps - 1
    print(f"[config 5] tracer {edge}^3 {'implicit' if implicit else 'explicit'} step: assemble {tot[0] / k:.3f} ms "
          f"({136 * n / (tot[0] / k) / 1e6:.0f} GB/s of 136 B/cell), solve {tot[1] / k:.2f} ms ({nit / k:.1f} BiCGSTAB its), update {tot[2] / k:.3f} ms "
          f"-> {n / (sum(tot) / k * 1e-3) / 1e6:.1f} MDOF/s")
    x = et.download(B.VEC_CUR)
    print(f"           mass fraction range [{x.min():.3e}, {x.max():.3e}]")
    et.close()
