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
# Dumux::ParMTJac / ParMTSOR / ParMTSSOR as AMG smoothers (test/linear/test_parallel_amg_smoothers.cc:25-40)
PARMT_SMOOTHERS = {"par_mt_jac": O.PARMT_JAC, "par_mt_sor": O.PARMT_SOR, "par_mt_ssor": O.PARMT_SSOR}


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


def _i32(v):
    return np.ascontiguousarray(v, dtype=np.int32)


class AmgOracle:
    """`cells`: the LOCAL box.  Block-decomposed run (dist_oracle.BoxRank): `layout` = RankLayout of the fine level, `part3` =
    ranks per axis, `gcells` = global cells; the hierarchy is then the GLOBAL one, every level decomposed like the grid with
    overlap 1 (aggregates never cross a processor boundary, see oracle.cpp orc_amg_galerkin), and the cycle contains the
    parallel pieces of dune-istl's overlapping AMG: smoother = BlockPreconditioner(local smoother; copyOwnerToAll), operator =
    local mv + project."""

    def __init__(self, cells, dim, b, rowptr, colidx, values, pre_steps=2, post_steps=2, damping=1.6, smoother=SMOOTHER_SSOR,
                 coarsest_cells=8, coarsest_steps=8, max_levels=15, layout=None, part3=None, gcells=None,
                 smoother_iterations=1, smoother_relaxation=1.0):
        self.b, self.dim = b, dim
        self.smoother, self.sm_it, self.sm_w = smoother, smoother_iterations, smoother_relaxation
        self.pre, self.post, self.damp, self.coarsest_steps = pre_steps, post_steps, damping, coarsest_steps
        self.levels = []
        c3 = tuple(cells) + (1,) * (3 - len(cells))
        lv = Level()
        lv.cells, lv.n = c3, int(np.prod(c3))
        lv.rowptr, lv.colidx = _i32(rowptr), _i32(colidx)
        lv.values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        lv.layout = layout if (layout is not None and layout.comm.nranks > 1) else None
        if lv.layout is None:
            part3, coord = (1, 1, 1), (0, 0, 0)
            lv.gcells = c3
            lv.ranges = [(0, c, 0, c) for c in c3]
        else:
            from dumux_b200 import problems
            coord = problems.rank_coord(part3, layout.comm.rank)
            lv.gcells = tuple(gcells) + (1,) * (3 - len(gcells))
            lv.ranges = list(layout.ranges)
        # owned sizes of ALL torus coordinates per axis (every rank computes the whole table: no communication)
        if lv.layout is None:
            sizes = [[c] for c in c3]
        else:
            from dumux_b200 import problems
            sizes = []
            for a in range(3):
                row = []
                for c in range(part3[a]):
                    _, _, b0, b1 = problems.axis_partition(lv.gcells[a], part3[a], c)
                    row.append(b1 - b0)
                sizes.append(row)
        lv.sizes = sizes
        self.levels.append(lv)
        while len(self.levels) < max_levels:
            f = self.levels[-1]
            if int(np.prod(f.gcells)) <= coarsest_cells:
                break
            csizes = [[(s + 1) // 2 for s in f.sizes[a]] if a < dim else list(f.sizes[a]) for a in range(3)]
            cg = tuple(sum(csizes[a]) for a in range(3))
            if cg == f.gcells:
                break
            lv = Level()
            lv.sizes, lv.gcells = csizes, cg
            lv.ranges = []
            for a in range(3):
                cb0 = sum(csizes[a][:coord[a]])
                cb1 = cb0 + csizes[a][coord[a]]
                lv.ranges.append((max(0, cb0 - 1), min(cg[a], cb1 + 1), cb0, cb1))
            lv.cells = tuple(r[1] - r[0] for r in lv.ranges)
            lv.n = int(np.prod(lv.cells))
            lv.layout = None
            if f.layout is not None:
                from oracle.dist_oracle import RankLayout
                lv.layout = RankLayout(f.layout.comm, part3, lv.ranges, b)
            lv.rowptr, lv.colidx = grid_pattern(lv.cells, dim)
            lv.values = np.zeros(int(lv.rowptr[-1]) * b * b)
            O.lib().orc_amg_galerkin(b, dim, _i32(f.cells), _i32(self._own(f)[0]), _i32(self._own(f)[1]), f.rowptr, f.values,
                                     _i32(lv.cells), _i32(self._own(lv)[0]), lv.rowptr, lv.values)
            self.levels.append(lv)
        self.status = 0
        for lv in self.levels:
            if smoother in PARMT_SMOOTHERS:
                continue
            if smoother == SMOOTHER_SSOR:
                lv.fac, st = O.ssor_factor(lv.n, b, lv.rowptr, lv.colidx, lv.values)
            else:
                lv.fac, st = O.ilu0_factor(lv.n, b, lv.rowptr, lv.colidx, lv.values)
            self.status = max(self.status, st)

    @staticmethod
    def _own(lv):
        """owned range per axis in LOCAL indices: (lo[3], hi[3])"""
        return [r[2] - r[0] for r in lv.ranges], [r[3] - r[0] for r in lv.ranges]

    # -- transfer operators on the box aggregates, children in lexicographic order --
    def _agg_maps(self, f, c):
        """per axis: local fine index -> local coarse index"""
        flo, fhi = self._own(f)
        clo, _ = self._own(c)
        maps = []
        for a in range(3):
            i = np.arange(f.cells[a])
            m = np.where(i < flo[a], clo[a] - 1, np.where(i >= fhi[a], clo[a] + ((fhi[a] - flo[a] + 1) >> 1), clo[a] + ((i - flo[a]) >> 1)))
            maps.append(m)
        return maps

    def _children(self, f, c):
        flo, fhi = self._own(f)
        clo, _ = self._own(c)
        first, count = [], []
        for a in range(3):
            I = np.arange(c.cells[a])
            chi = clo[a] + ((fhi[a] - flo[a] + 1) >> 1)
            c0 = np.where(I < clo[a], flo[a] - 1, np.where(I >= chi, fhi[a], flo[a] + 2 * (I - clo[a])))
            c1 = np.where(I < clo[a], flo[a], np.where(I >= chi, fhi[a] + 1, np.minimum(flo[a] + 2 * (I - clo[a]) + 2, fhi[a])))
            first.append(c0); count.append(c1 - c0)
        for dz in range(2):
            for dy in range(2):
                for dx in range(2):
                    ok = (dz < count[2])[:, None, None] & (dy < count[1])[None, :, None] & (dx < count[0])[None, None, :]
                    kk = np.clip(first[2] + dz, 0, f.cells[2] - 1)
                    jj = np.clip(first[1] + dy, 0, f.cells[1] - 1)
                    ii = np.clip(first[0] + dx, 0, f.cells[0] - 1)
                    yield (kk, jj, ii), ok

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
        mi, mj, mk = self._agg_maps(f, c)
        return (self.damp * g[np.ix_(mk, mj, mi)]).reshape(-1)

    def _mv(self, lv, u):
        t = O.spmv(lv.n, self.b, lv.rowptr, lv.colidx, lv.values, u)
        if lv.layout is not None:
            t[~lv.layout.owner] = 0.0                      # OverlappingSchwarzOperator: project
        return t

    def _smooth_step(self, lv, x, r, first, need_defect):
        if self.smoother in PARMT_SMOOTHERS:
            u = O.parmt_apply(PARMT_SMOOTHERS[self.smoother], lv.n, self.b, lv.rowptr, lv.colidx, lv.values, r, self.sm_it, self.sm_w)
        else:
            u = O.ilu0_apply(lv.n, self.b, lv.rowptr, lv.colidx, lv.fac, r)
        if lv.layout is not None:
            lv.layout.copy_owner_to_all(u)                 # BlockPreconditioner::apply
        x = u if first else x + u
        if need_defect:
            r = r - self._mv(lv, u)
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
            r = r - self._mv(lv, u)
        for s in range(self.post):
            x, r = self._smooth_step(lv, x, r, False, s + 1 < self.post)
        return x

    def apply(self, d):
        """v = AMG(d): one V-cycle from v = 0 (block-decomposed: v comes back consistent on the overlap)"""
        return self.cycle(0, d)
