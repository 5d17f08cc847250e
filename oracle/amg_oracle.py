"""Aggregation AMG on the structured grid hierarchy -- CPU restatement of dumux_b200/csrc/amg.cu.  TEST INFRASTRUCTURE ONLY.

The cycle is dune-istl's default AMG cycle (AMGBiCGSTABIstlSolver, dumux/linear/istlsolvers.hh:716-757 -> Dune::AMGCreator):
V-cycle, preSteps = postSteps = 2, prolongationDampingFactor 1.6, smoother SeqSSOR (or SeqILU), a smoothing step being
"update = 0; smoother.apply(update, defect); lhs += update; defect -= A update" [DUNE-ext, paamg/amg.hh].  dune's aggregation
heuristic is NOT restated (it is a graph algorithm that lives in dune-istl, absent here): the aggregates are the 2x2x2 cell
blocks of the structured box, which keeps every coarse matrix on the 7-point pattern; the hierarchy ends at <= coarsest_cells
cells, where coarsest_steps smoothing steps replace dune's direct coarse solve.  Every sum is taken in the order the CUDA
kernels use (children of an aggregate in lexicographic order), so the device reproduces this cycle bit for bit.
"""
from __future__ import annotations

import numpy as np

from oracle import oracle_py as O

SMOOTHER_SSOR, SMOOTHER_ILU0 = "ssor", "ilu"


def grid_pattern(cells, dim):
    """BCRS pattern of the 7-point stencil on `cells` (x fastest), columns ascending -- what dmx_grid_structured builds"""
    nx, ny, nz = cells
    n = nx * ny * nz
    idx = np.arange(n)
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    cols = [(np.where((k > 0) & (dim > 2), idx - nx * ny, -1)), (np.where((j > 0) & (dim > 1), idx - nx, -1)), (np.where(i > 0, idx - 1, -1)), idx,
            (np.where(i + 1 < nx, idx + 1, -1)), (np.where((j + 1 < ny) & (dim > 1), idx + nx, -1)), (np.where((k + 1 < nz) & (dim > 2), idx + nx * ny, -1))]
    c = np.stack(cols, axis=1)
    mask = c >= 0
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(mask.sum(axis=1))
    return rowptr, c[mask].astype(np.int32)


class Level:
    pass


class AmgOracle:
    def __init__(self, cells, dim, b, rowptr, colidx, values, pre_steps=2, post_steps=2, damping=1.6, smoother=SMOOTHER_SSOR,
                 coarsest_cells=8, coarsest_steps=8, max_levels=15):
        self.b, self.dim = b, dim
        self.pre, self.post, self.damp, self.coarsest_steps = pre_steps, post_steps, damping, coarsest_steps
        self.levels = []
        c3 = tuple(cells) + (1,) * (3 - len(cells))
        lv = Level()
        lv.cells, lv.n = c3, int(np.prod(c3))
        lv.rowptr, lv.colidx = np.ascontiguousarray(rowptr, dtype=np.int32), np.ascontiguousarray(colidx, dtype=np.int32)
        lv.values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        self.levels.append(lv)
        while len(self.levels) < max_levels:
            f = self.levels[-1]
            if f.n <= coarsest_cells or all(c == 1 for c in f.cells[:dim]):
                break
            cc = tuple((f.cells[a] + 1) // 2 if a < dim else 1 for a in range(3))
            lv = Level()
            lv.cells, lv.n = cc, int(np.prod(cc))
            lv.rowptr, lv.colidx = grid_pattern(cc, dim)
            lv.values = np.zeros(int(lv.rowptr[-1]) * b * b)
            O.lib().orc_amg_galerkin(b, dim, np.ascontiguousarray(f.cells, dtype=np.int32), f.rowptr, f.values,
                                     np.ascontiguousarray(cc, dtype=np.int32), lv.rowptr, lv.values)
            self.levels.append(lv)
        self.status = 0
        for lv in self.levels:
            if smoother == SMOOTHER_SSOR:
                lv.fac, st = O.ssor_factor(lv.n, b, lv.rowptr, lv.colidx, lv.values)
            else:
                lv.fac, st = O.ilu0_factor(lv.n, b, lv.rowptr, lv.colidx, lv.values)
            self.status = max(self.status, st)

    # -- transfer operators on 2x2x2 box aggregates, children in lexicographic order --
    def _children(self, f, c):
        fx, fy, fz = f.cells
        cx, cy, cz = c.cells
        for dz in range(2):
            for dy in range(2):
                for dx in range(2):
                    ii, jj, kk = np.arange(cx) * 2 + dx, np.arange(cy) * 2 + dy, np.arange(cz) * 2 + dz
                    ok = (kk < fz)[:, None, None] & (jj < fy)[None, :, None] & (ii < fx)[None, None, :]
                    yield (np.minimum(kk, fz - 1), np.minimum(jj, fy - 1), np.minimum(ii, fx - 1)), ok

    def restrict(self, l, r):
        f, c = self.levels[l], self.levels[l + 1]
        rf = r.reshape(f.cells[2], f.cells[1], f.cells[0], self.b)
        s = np.zeros((c.cells[2], c.cells[1], c.cells[0], self.b))
        for (kk, jj, ii), ok in self._children(f, c):
            v = rf[np.ix_(kk, jj, ii)]
            s = np.where(ok[..., None], s + v, s)
        return s.reshape(-1)

    def prolong(self, l, xc):
        f, c = self.levels[l], self.levels[l + 1]
        g = xc.reshape(c.cells[2], c.cells[1], c.cells[0], self.b)
        kk, jj, ii = np.arange(f.cells[2]) // 2, np.arange(f.cells[1]) // 2, np.arange(f.cells[0]) // 2
        return (self.damp * g[np.ix_(kk, jj, ii)]).reshape(-1)

    def _smooth_step(self, lv, x, r, first, need_defect):
        u = O.ilu0_apply(lv.n, self.b, lv.rowptr, lv.colidx, lv.fac, r)
        x = u if first else x + u
        if need_defect:
            r = r - O.spmv(lv.n, self.b, lv.rowptr, lv.colidx, lv.values, u)
        return x, r

    def cycle(self, l, d):
        lv = self.levels[l]
        r = np.array(d, dtype=np.float64)
        x = np.zeros_like(r)
        if l + 1 == len(self.levels):
            ns = max(1, self.coarsest_steps)
            for s in range(ns):
                x, r = self._smooth_step(lv, x, r, s == 0, s + 1 < ns)
            return x
        for s in range(self.pre):
            x, r = self._smooth_step(lv, x, r, s == 0, True)
        xc = self.cycle(l + 1, self.restrict(l, r))
        u = self.prolong(l, xc)
        x = u if self.pre == 0 else x + u
        if self.post > 0:
            r = r - O.spmv(lv.n, self.b, lv.rowptr, lv.colidx, lv.values, u)
        for s in range(self.post):
            x, r = self._smooth_step(lv, x, r, False, s + 1 < self.post)
        return x

    def apply(self, d):
        """v = AMG(d): one V-cycle from v = 0"""
        return self.cycle(0, d)
