"""Block-decomposed (overlapping Schwarz) restatement of the Newton step on the CPU oracle.  TEST INFRASTRUCTURE ONLY.

What DuMux does for a cell-centred scheme on P MPI ranks (SURVEY 2.2, 8e, Appendix A "Overlapping variant"):
  * YaspGrid cuts the structured box into px x py x pz blocks (Grid.Partitioning, io/grid/gridmanager_yasp.hh:194-203; slabs
    "1 1 P" by default here) and adds `Grid.Overlap 1` layers; every rank
    assembles the rows of ALL its cells (interior + overlap); the outer face of an overlap cell carries no scvf
    (discretization/cellcentered/tpfa/fvgridgeometry.hh:272-320), so those rows are incomplete;
  * linear/linearsolvertraits.hh:79-91 + linear/istlsolvers.hh:550-566 wire dune-istl's
    OverlappingSchwarzOperator (local A.mv, then project = zero the non-owner entries),
    OverlappingSchwarzScalarProduct (owner-masked dot + global sum) and
    BlockPreconditioner (pre: copyOwnerToAll(x); apply: local SeqILU, then copyOwnerToAll(v)) around BiCGSTABSolver.

The algorithm below is written SPMD (one call per rank) against a tiny communicator interface with two implementations:
`ThreadComm` (P python threads in one process: the reference run) and `TorchComm` (torch.distributed, gloo or nccl: the
world_size-2 CPU tests).  The GPU path (dumux_b200/csrc/dist.cu + linalg.cu) is compared against it.
"""
from __future__ import annotations

import math
import threading

import numpy as np

from dumux_b200 import problems
from oracle import oracle_py as O


# ----------------------------------------------------------------------------------------------------------
# The summation tree of the CUDA reductions (dumux_b200/csrc/linalg.cu: dot_kernel / axpy*_norm* / residual_init ->
# block_reduce -> final_reduce_kernel), restated so that a scalar product here returns the SAME BITS as on the device:
# 1184 x 256 threads each add their grid-stride elements in order, a warp adds lane l and l+16, l+8, l+4, l+2, l+1, the
# 8 warp sums of a block go through the same tree (padded with zeros), and the 1184 block partials are summed by 256
# threads in stride order followed by one more block tree.  dune-istl's SeqScalarProduct sums sequentially instead
# (`BoxRank(gpu_reduction=False)`, the default); the two differ in the last bits of every dot product, which BiCGSTAB
# amplifies into iteration counts that wander by a few -- with this tree the device and the oracle run the identical
# iteration, so the counts can be compared exactly.
# ----------------------------------------------------------------------------------------------------------
RED_BLOCKS, RED_THREADS = 1184, 256


def _tree32(v):
    """lane 0 of `for o in 16,8,4,2,1: v += shfl_down(v, o)` over the last axis (length 32)"""
    for o in (16, 8, 4, 2, 1):
        v = v[..., :o] + v[..., o:2 * o]
    return v[..., 0]


def _block_tree(v):
    """block_reduce<false> of linalg.cu: v[..., 256] -> 8 warp sums -> one more warp tree with lanes 8..31 = 0"""
    w = _tree32(v.reshape(v.shape[:-1] + (RED_THREADS // 32, 32)))
    pad = np.zeros(w.shape[:-1] + (32,))
    pad[..., :w.shape[-1]] = w
    return _tree32(pad)


def gpu_sum(terms):
    """sum of `terms` (one per vector entry, non-owner entries already zero) in the device's summation order"""
    G = RED_BLOCKS * RED_THREADS
    L = terms.size
    nch = max(1, -(-L // G))
    pad = np.zeros(nch * G)
    pad[:L] = terms
    acc = np.zeros(G)
    for k in range(nch):                                      # each thread: s += term, grid-stride order
        acc = acc + pad[k * G:(k + 1) * G]
    partials = _block_tree(acc.reshape(RED_BLOCKS, RED_THREADS))
    nch2 = -(-RED_BLOCKS // RED_THREADS)
    pad2 = np.zeros(nch2 * RED_THREADS)
    pad2[:RED_BLOCKS] = partials
    acc2 = np.zeros(RED_THREADS)
    for k in range(nch2):                                     # final_reduce_kernel: thread t adds partials[t], [t+256], ...
        acc2 = acc2 + pad2[k * RED_THREADS:(k + 1) * RED_THREADS]
    return float(_block_tree(acc2))


# ----------------------------------------------------------------------------------------------------------
# communicators
# ----------------------------------------------------------------------------------------------------------
class ThreadComm:
    """P ranks as threads of one process; collectives through a barrier and shared slots (deterministic rank order)."""

    class Shared:
        def __init__(self, nranks):
            self.nranks = nranks
            self.barrier = threading.Barrier(nranks)
            self.slots = [None] * nranks
            self.mail = {}

    def __init__(self, shared, rank):
        self.s, self.rank, self.nranks = shared, rank, shared.nranks

    def allreduce(self, value, op="sum"):
        self.s.slots[self.rank] = value
        self.s.barrier.wait()
        vals = list(self.s.slots)
        self.s.barrier.wait()
        if op == "sum":
            acc = vals[0]
            for v in vals[1:]:
                acc = acc + v
            return acc
        return max(vals) if op == "max" else min(vals)

    def exchange(self, sends):
        """`sends`: {neighbour rank: array}; returns {neighbour rank: the array that neighbour sent to me}."""
        for dst, a in sends.items():
            self.s.mail[(self.rank, dst)] = a.copy()
        self.s.barrier.wait()
        out = {src: self.s.mail[(src, self.rank)] for src in sends}
        self.s.barrier.wait()
        return out


class TorchComm:
    """torch.distributed backend (gloo on CPU)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.nranks = dist.get_rank(), dist.get_world_size()

    def allreduce(self, value, op="sum"):
        import torch
        t = torch.tensor(np.atleast_1d(np.asarray(value, dtype=np.float64)))
        self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN}[op])
        out = t.numpy()
        return float(out[0]) if np.ndim(value) == 0 else out

    def exchange(self, sends):
        import torch
        reqs, recv = [], {}
        for nb in sorted(sends):
            recv[nb] = torch.empty(sends[nb].size, dtype=torch.float64)
            reqs.append(self.dist.isend(torch.from_numpy(np.ascontiguousarray(sends[nb])), nb))
            reqs.append(self.dist.irecv(recv[nb], nb))
        for r in reqs:
            r.wait()
        return {nb: t.numpy() for nb, t in recv.items()}


# ----------------------------------------------------------------------------------------------------------
# one rank
# ----------------------------------------------------------------------------------------------------------
class RankLayout:
    """Owner mask and copyOwnerToAll neighbours of one rank's box: per axis (lo, hi, b0, b1) = local box incl. overlap and
    owned range in GLOBAL indices (3 axes; unused axes (0, 1, 0, 1)), `part3` ranks per axis, block size b."""

    def __init__(self, comm, part3, ranges3, b):
        import itertools
        self.comm, self.b = comm, b
        self.ranges = list(ranges3)
        lc3 = tuple(r[1] - r[0] for r in ranges3)
        self.cells3 = lc3
        self.off3 = tuple(r[0] for r in ranges3)
        self.n = int(np.prod(lc3))
        self.shape = (lc3[2], lc3[1], lc3[0], b)               # local block vector viewed [z, y, x, eq]
        own = np.ones(self.shape[:3], dtype=bool)
        self.own_rng = []
        for a in range(3):
            lo, hi, b0, b1 = ranges3[a]
            self.own_rng.append((b0 - lo, b1 - lo))
            idx = np.arange(lc3[a])
            ok = (idx >= b0 - lo) & (idx < b1 - lo)
            shp = [1, 1, 1]
            shp[2 - a] = -1
            own &= ok.reshape(shp)
        self.owner_cells = own.reshape(-1)
        self.owner = np.repeat(self.owner_cells, b)               # per scalar dof
        coord = problems.rank_coord(part3, comm.rank)
        self.neighbours = []
        for dz, dy, dx in itertools.product((-1, 0, 1), repeat=3):
            d = (dx, dy, dz)
            if d == (0, 0, 0):
                continue
            c = [coord[a] + d[a] for a in range(3)]
            if any(c[a] < 0 or c[a] >= part3[a] for a in range(3)):
                continue
            nb = c[0] + part3[0] * (c[1] + part3[1] * c[2])
            snd, rcv = [], []
            for a in range(3):
                o0, o1 = self.own_rng[a]
                if d[a] < 0:
                    snd.append(slice(o0, o0 + 1)); rcv.append(slice(o0 - 1, o0))
                elif d[a] > 0:
                    snd.append(slice(o1 - 1, o1)); rcv.append(slice(o1, o1 + 1))
                else:
                    snd.append(slice(o0, o1)); rcv.append(slice(o0, o1))
            self.neighbours.append((nb, tuple(reversed(snd)), tuple(reversed(rcv))))

    # copyOwnerToAll with overlap 1: owned cells inside a neighbour's overlap go out, my overlap cells come in from their owner
    def copy_owner_to_all(self, v):
        if not self.neighbours:
            return
        g = v.reshape(self.shape)
        got = self.comm.exchange({nb: np.ascontiguousarray(g[snd]).reshape(-1) for nb, snd, _ in self.neighbours})
        for nb, _, rcv in self.neighbours:
            g[rcv] = got[nb].reshape(g[rcv].shape)


class BoxRank:
    """Local problem of one rank: block [lo, hi) per axis incl. overlap, owned (interior) range [b0, b1) per axis.
    `part` = Grid.Partitioning (ranks per axis); None = slabs along the last axis.  `make_spec(box)` builds the box-local
    ProblemSpec (box = per-axis (lo, hi), None for the single-domain run)."""

    def __init__(self, make_spec, cells, comm, part=None, num_threads=0, gpu_reduction=False):
        import copy
        import itertools
        self.comm = comm
        self.gpu_reduction = gpu_reduction
        self.amg_params = None      # keyword arguments of oracle.amg_oracle.AmgOracle (precond="amg")
        dim = len(cells)
        self.cells = tuple(cells)
        self.part = tuple(part) if part is not None else problems.default_partitioning(dim, comm.nranks)
        assert int(np.prod(self.part)) == comm.nranks
        self.ranges = problems.box_partition(cells, self.part, comm.rank)          # per axis (lo, hi, b0, b1)
        self.box = [(r[0], r[1]) for r in self.ranges]
        # slab view of the last axis (kept for the slab tests)
        self.lo, self.hi, self.b0, self.b1 = self.ranges[-1]
        self.spec = make_spec(self.box) if comm.nranks > 1 else make_spec(None)
        spec = self.spec
        # the local oracle works on the local box: cut the geometry, mark processor boundaries as "no scvf"
        loc = copy.copy(spec)
        loc.cells = tuple(hi - lo for lo, hi in self.box)
        nodes = problems.node_coords(cells, spec.lower, spec.upper)
        loc.node_coords = [nodes[a][self.box[a][0]:self.box[a][1] + 1] for a in range(dim)]
        loc.lower = tuple(float(loc.node_coords[a][0]) for a in range(dim))
        loc.upper = tuple(float(loc.node_coords[a][-1]) for a in range(dim))
        loc.bc_type = dict(spec.bc_type)
        loc.bc_values = dict(spec.bc_values)
        for a in range(dim):
            nf = int(np.prod([loc.cells[d] for d in range(dim) if d != a]))
            if self.box[a][0] > 0:
                loc.bc_type[2 * a] = np.full(nf, problems.BC_NONE, dtype=np.int32)
                loc.bc_values[2 * a] = np.zeros((nf, spec.num_eq))
            if self.box[a][1] < cells[a]:
                loc.bc_type[2 * a + 1] = np.full(nf, problems.BC_NONE, dtype=np.int32)
                loc.bc_values[2 * a + 1] = np.zeros((nf, spec.num_eq))
        loc.slab = None
        loc.box = None
        self.local = loc
        self.o = O.Oracle(loc, num_threads=num_threads)
        self.b = self.o.b
        self.n = self.o.n
        part3 = tuple(self.part) + (1,) * (3 - dim)
        self.layout = RankLayout(comm, part3, list(self.ranges) + [(0, 1, 0, 1)] * (3 - dim), self.b)
        self.part3 = part3
        self.shape, self.own_rng, self.owner, self.neighbours = self.layout.shape, self.layout.own_rng, self.layout.owner, self.layout.neighbours
        # index lists of the owner / non-owner entries and gather buffers: the same compact arrays a boolean mask would produce,
        # without rebuilding them (and holding the interpreter lock) in every scalar product of the Krylov loops
        if comm.nranks > 1:
            self._oidx = np.flatnonzero(self.owner)
            self._nidx = np.flatnonzero(~self.owner)
            self._ga, self._gb = np.empty(self._oidx.size), np.empty(self._oidx.size)
        else:
            self._oidx = self._nidx = None

    def copy_owner_to_all(self, v):
        self.layout.copy_owner_to_all(v)

    def gather_box(self):
        """(global slices [z, y, x] of my owned block, local slices of it)"""
        dim = len(self.cells)
        gs, ls = [], []
        for a in range(3):
            if a < dim:
                lo, hi, b0, b1 = self.ranges[a]
                gs.append(slice(b0, b1)); ls.append(slice(b0 - lo, b1 - lo))
            else:
                gs.append(slice(0, 1)); ls.append(slice(0, 1))
        return tuple(reversed(gs)), tuple(reversed(ls))

    def dot(self, a, b):
        if self.gpu_reduction:
            return self.comm.allreduce(gpu_sum(np.where(self.owner, a * b, 0.0)), "sum")
        if self._oidx is None:
            return self.comm.allreduce(float(np.dot(a, b)), "sum")
        a.take(self._oidx, out=self._ga)
        if b is a:
            return self.comm.allreduce(float(np.dot(self._ga, self._ga)), "sum")
        b.take(self._oidx, out=self._gb)
        return self.comm.allreduce(float(np.dot(self._ga, self._gb)), "sum")

    def project(self, y):
        """OverlappingSchwarzOperator: zero the non-owner entries"""
        if self._nidx is not None:
            y[self._nidx] = 0.0
        return y

    def apply_operator(self, jac, x):
        return self.project(O.spmv(self.n, self.b, self.o.rowptr, self.o.colidx, jac, x))

    def local_preconditioner(self, jac, precond="ilu0", iterations=1, relaxation=1.0):
        """The sequential preconditioner of this rank's (interior + overlap) matrix: d -> M^-1 d, or None if the set-up fails
        on any rank.  "ilu0" = SeqILU(0), "ssor" = SeqSSOR, "par_mt_jac" / "par_mt_sor" / "par_mt_ssor" = Dumux::ParMT*."""
        n, b, rp, ci = self.n, self.b, self.o.rowptr, self.o.colidx
        if precond == "ilu0":
            ilu, st = O.ilu0_factor(n, b, rp, ci, jac)
            st = int(self.comm.allreduce(float(st), "max"))
            if st != 0:
                return None
            return lambda d: O.ilu0_apply(n, b, rp, ci, ilu, d)
        if precond == "ssor":
            return lambda d: O.ssor_apply(n, b, rp, ci, jac, d)
        if precond == "amg":
            # the GLOBAL hierarchy, block-decomposed like the grid: its cycle contains the owner -> copy exchanges itself
            from oracle.amg_oracle import AmgOracle
            amg = AmgOracle(self.local.cells, len(self.local.cells), b, rp, ci, jac, layout=self.layout, part3=self.part3,
                            gcells=self.cells, **(self.amg_params or {}))
            st = int(self.comm.allreduce(float(amg.status), "max"))
            return None if st != 0 else amg.apply
        kind = {"par_mt_jac": O.PARMT_JAC, "par_mt_sor": O.PARMT_SOR, "par_mt_ssor": O.PARMT_SSOR}[precond]
        return lambda d: O.parmt_apply(kind, n, b, rp, ci, jac, d, iterations, relaxation)

    def bicgstab(self, jac, rhs, reduction=1e-6, maxit=250, precond="ilu0", iterations=1, relaxation=1.0):
        """Dune::BiCGSTABSolver::apply with the overlapping-Schwarz operator / scalar product / BlockPreconditioner; x0 = 0."""
        local = self.local_preconditioner(jac, precond, iterations, relaxation)
        if local is None:
            return np.zeros_like(rhs), 2, 0, 1.0

        def prec(d):
            v = local(d)
            self.copy_owner_to_all(v)
            return v

        x = np.zeros_like(rhs)
        self.copy_owner_to_all(x)                                    # BlockPreconditioner::pre
        r = rhs - O.spmv(self.n, self.b, self.o.rowptr, self.o.colidx, jac, x)
        self.project(r)                                              # applyscaleadd(-1, x, r) projects r
        rt = r.copy()
        norm0 = math.sqrt(self.dot(r, r))
        norm = norm0
        conv = lambda nrm: nrm < reduction * norm0 or nrm < 1e-30
        if not math.isfinite(norm0):
            return x, 3, 0, 1.0
        if conv(norm0):
            return x, 0, 0, (1.0 if norm0 > 0 else 0.0)
        p = np.zeros_like(rhs)
        v = np.zeros_like(rhs)
        tmp = np.empty_like(rhs)                                     # the updates below are the same operations in place
        rho = alpha = omega = 1.0
        it = 0.5
        status = 1
        while it < maxit:
            rho_new = self.dot(rt, r)
            if abs(rho) <= 1e-80 or abs(omega) <= 1e-80:
                status = 2
                break
            if it < 1:
                p = r.copy()
            else:
                beta = (rho_new / rho) * (alpha / omega)
                np.multiply(v, -omega, out=tmp)                      # p = (p + (-omega) v) beta + r
                np.add(p, tmp, out=p)
                np.multiply(p, beta, out=p)
                np.add(p, r, out=p)
            y = prec(p)
            v = self.apply_operator(jac, y)
            h = self.dot(rt, v)
            if abs(h) < 1e-80:
                status = 2
                break
            alpha = rho_new / h
            np.multiply(y, alpha, out=tmp)
            np.add(x, tmp, out=x)                                    # x += alpha y
            np.multiply(v, -alpha, out=tmp)
            np.add(r, tmp, out=r)                                    # r += (-alpha) v
            norm = math.sqrt(self.dot(r, r))
            if not math.isfinite(norm):
                status = 3
                break
            if conv(norm):
                status = 0
                break
            it += 0.5
            y = prec(r)
            t = self.apply_operator(jac, y)
            omega = self.dot(t, r) / self.dot(t, t)
            np.multiply(y, omega, out=tmp)
            np.add(x, tmp, out=x)                                    # x += omega y
            np.multiply(t, -omega, out=tmp)
            np.add(r, tmp, out=r)                                    # r += (-omega) t
            rho = rho_new
            norm = math.sqrt(self.dot(r, r))
            if not math.isfinite(norm):
                status = 3
                break
            if conv(norm):
                status = 0
                break
            it += 0.5
        return x, status, int(math.ceil(min(it, maxit))), (norm / norm0 if norm0 > 0 else 0.0)

    def cg(self, jac, rhs, reduction=1e-6, maxit=250, precond="ssor", iterations=1, relaxation=1.0):
        """Dune::CGSolver::apply (see oracle.cpp cgSolve) with the overlapping-Schwarz operator / scalar product /
        BlockPreconditioner; x0 = 0.  Returns (x, status, iterations, achieved reduction)."""
        local = self.local_preconditioner(jac, precond, iterations, relaxation)
        if local is None:
            return np.zeros_like(rhs), 2, 0, 1.0

        def prec(d):
            v = local(d)
            self.copy_owner_to_all(v)
            return v

        x = np.zeros_like(rhs)
        self.copy_owner_to_all(x)
        b = rhs - O.spmv(self.n, self.b, self.o.rowptr, self.o.colidx, jac, x)
        self.project(b)
        def0 = math.sqrt(self.dot(b, b))
        if not math.isfinite(def0):
            return x, 3, 0, 1.0
        conv = lambda nrm: nrm < reduction * def0 or nrm < 1e-30
        if conv(def0):
            return x, 0, 0, (1.0 if def0 > 0 else 0.0)
        p = prec(b)
        rholast = self.dot(p, b)
        deff, status, i = def0, 1, 1
        while i <= maxit:
            q = self.apply_operator(jac, p)
            alpha = self.dot(p, q)
            lam = rholast / alpha
            x += lam * p
            b += (-lam) * q
            deff = math.sqrt(self.dot(b, b))
            its = i
            if not math.isfinite(deff):
                status = 3
                break
            if conv(deff):
                status = 0
                break
            q = prec(b)
            rho = self.dot(q, b)
            beta = rho / rholast
            p = p * beta
            p += q
            rholast = rho
            i += 1
        else:
            its = maxit
        return x, status, its, (deff / def0 if def0 > 0 else 0.0)

    def gmres(self, jac, rhs, reduction=1e-6, maxit=250, restart=10):
        """Dune::RestartedGMResSolver::apply (left preconditioned, see oracle.cpp restartedGmres) with the overlapping-Schwarz
        operator / scalar product / BlockPreconditioner; x0 = 0.  Returns (x, status, iterations, achieved reduction)."""
        ilu, st = O.ilu0_factor(self.n, self.b, self.o.rowptr, self.o.colidx, jac)
        st = int(self.comm.allreduce(float(st), "max"))
        if st != 0:
            return np.zeros_like(rhs), 2, 0, 1.0

        def prec(d):
            v = O.ilu0_apply(self.n, self.b, self.o.rowptr, self.o.colidx, ilu, d)
            self.copy_owner_to_all(v)
            return v

        def rotation(dx, dy):
            eps = 1e-15
            ndx, ndy = abs(dx), abs(dy)
            if ndy < eps:
                return 1.0, 0.0
            if ndx < eps:
                return 0.0, 1.0
            temp = min(ndx, ndy) / max(ndx, ndy)
            if ndy > ndx:
                return 1.0 / math.sqrt(1.0 + temp * temp) * temp, 1.0 / math.sqrt(1.0 + temp * temp) * dx * dy / ndx / ndy
            return 1.0 / math.sqrt(1.0 + temp * temp), 1.0 / math.sqrt(1.0 + temp * temp) * dy / dx

        m = restart
        x = np.zeros_like(rhs)
        self.copy_owner_to_all(x)                                    # BlockPreconditioner::pre

        def defect():
            b = rhs - O.spmv(self.n, self.b, self.o.rowptr, self.o.colidx, jac, x)
            self.project(b)                                          # applyscaleadd projects
            v0 = prec(b)
            return v0, math.sqrt(self.dot(v0, v0))

        v = [None] * (m + 1)
        v[0], norm = defect()
        norm0 = norm
        if not math.isfinite(norm0):
            return x, 3, 0, 1.0
        conv = lambda nrm: nrm < reduction * norm0 or nrm < 1e-30
        if conv(norm0):
            return x, 0, 0, (1.0 if norm0 > 0 else 0.0)
        H = np.zeros((m + 1, m + 1))
        cs, sn, s = np.zeros(m), np.zeros(m), np.zeros(m + 1)
        j, converged, status, its = 1, False, 1, 0
        while j <= maxit and not converged:
            v[0] = v[0] * (0.0 if norm == 0.0 else 1.0 / norm)
            s[:] = 0.0
            s[0] = norm
            i = 0
            while i < m and j <= maxit and not converged:
                w = prec(self.apply_operator(jac, v[i]))
                for k in range(i + 1):
                    H[k, i] = self.dot(v[k], w)
                    w = w + (-H[k, i]) * v[k]
                H[i + 1, i] = math.sqrt(self.dot(w, w))
                its = j
                if not math.isfinite(H[i + 1, i]):
                    return x, 3, its, 1.0
                if abs(H[i + 1, i]) < 1e-80:
                    return x, 2, its, norm / norm0
                v[i + 1] = w * (0.0 if norm == 0.0 else 1.0 / H[i + 1, i])
                for k in range(i):
                    t = cs[k] * H[k, i] + sn[k] * H[k + 1, i]
                    H[k + 1, i] = -sn[k] * H[k, i] + cs[k] * H[k + 1, i]
                    H[k, i] = t
                cs[i], sn[i] = rotation(H[i, i], H[i + 1, i])
                t = cs[i] * H[i, i] + sn[i] * H[i + 1, i]
                H[i + 1, i] = -sn[i] * H[i, i] + cs[i] * H[i + 1, i]
                H[i, i] = t
                t = cs[i] * s[i] + sn[i] * s[i + 1]
                s[i + 1] = -sn[i] * s[i] + cs[i] * s[i + 1]
                s[i] = t
                norm = abs(s[i + 1])
                if conv(norm):
                    converged, status = True, 0
                i += 1
                j += 1
            y = s.copy()
            w = np.zeros_like(rhs)
            for a in range(i - 1, -1, -1):
                r = s[a]
                for c in range(a + 1, i):
                    r -= H[a, c] * y[c]
                y[a] = 0.0 if r == 0.0 else r / H[a, a]
                w = w + y[a] * v[a]
            x = x + w
            if not converged and j < maxit:
                v[0], norm = defect()
        return x, status, its, (norm / norm0 if norm0 > 0 else 0.0)

    def newton(self, u0, prev, lin_reduction=1e-6, lin_maxit=250, max_rel_shift=1e-8, min_steps=2, max_steps=18, precond="ilu0"):
        """NewtonSolver::solveImpl_ (nonlinear/newtonsolver.hh:976-1072) on the local slab; u0/prev are LOCAL arrays."""
        u = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1).copy()
        prev = np.ascontiguousarray(prev, dtype=np.float64).reshape(-1)
        steps, shift, last_shift, converged = 0, 0.0, 0.0, False
        lin_its = []
        while True:
            if steps >= min_steps:
                if converged:
                    break
                if steps >= max_steps and not (shift * 4.0 < last_shift):
                    break
            last_shift = shift
            res, jac = self.o.assemble(u, prev)
            dx, st, its, red = self.bicgstab(jac, res, lin_reduction, lin_maxit, precond=precond)
            lin_its.append(its)
            if st != 0:
                return u, st, steps, lin_its
            u_new = u + (-1.0) * dx
            sh = np.abs(u_new - u) / np.maximum(1.0, np.abs(u_new + u) * 0.5)
            shift = self.comm.allreduce(float(sh[self.owner].max()), "max")
            u = u_new
            steps += 1
            converged = shift <= max_rel_shift
        return u, (0 if converged else 1), steps, lin_its


SlabRank = BoxRank      # the slab decomposition is the default partitioning of BoxRank


def single_rank(spec, gpu_reduction=True, num_threads=0):
    """The one-rank case as a BoxRank: the same BiCGSTAB / Newton loops as the multi-rank reference, with the device's summation
    tree for the scalar products by default -- what a single-GPU run is compared with when iteration COUNTS must match exactly."""
    return BoxRank(lambda box: spec, spec.cells, ThreadComm(ThreadComm.Shared(1), 0), None, num_threads, gpu_reduction)


def run_threads(make_spec, cells, nranks, fn, part=None, num_threads=0, gpu_reduction=False):
    """Runs fn(BoxRank) on `nranks` threads; returns the per-rank results."""
    import sys
    shared = ThreadComm.Shared(nranks)
    out = [None] * nranks
    err = []
    # the ranks hand over to each other at every collective (two barrier waits per all-reduce): with CPython's default 5 ms
    # switch interval a rank that arrives at a barrier can wait that long for the interpreter lock
    old_interval = sys.getswitchinterval()
    sys.setswitchinterval(2e-5)

    def work(r):
        try:
            out[r] = fn(BoxRank(make_spec, cells, ThreadComm(shared, r), part, num_threads, gpu_reduction))
        except BaseException as e:       # noqa: BLE001
            err.append(e)
            shared.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    sys.setswitchinterval(old_interval)
    if err:
        raise err[0]
    return out


def native_schwarz_bicgstab(cells, part, b, systems, reduction=1e-6, maxit=250):
    """dist_oracle.BoxRank.bicgstab for ALL ranks at once in native code (oracle.cpp orc_schwarz_ilu0_bicgstab: one OpenMP
    thread per rank, no interpreter in the loop) -- what bench.py's CPU arms time.  `systems[r]` = (rowptr, colidx, values, rhs)
    of rank r's local box (overlap rows included); returns (x per rank, status, iterations, achieved reduction, seconds)."""
    import ctypes as C
    dim = len(cells)
    part = tuple(part) if part is not None else problems.default_partitioning(dim, len(systems))
    nranks = int(np.prod(part))
    assert nranks == len(systems)
    part3 = tuple(part) + (1,) * (3 - dim)
    c3 = tuple(cells) + (1,) * (3 - dim)
    rngs = [problems.box_partition(cells, part, r) + [(0, 1, 0, 1)] * (3 - dim) for r in range(nranks)]
    # owner coordinate of every global index per axis
    owner_coord = []
    for a in range(3):
        oc = np.zeros(c3[a], dtype=np.int64)
        for c in range(part3[a]):
            lo, hi, b0, b1 = problems.axis_partition(c3[a], part3[a], c)
            oc[b0:b1] = c
        owner_coord.append(oc)
    owners, dsts, src_ranks, src_idxs, ns = [], [], [], [], []
    for r in range(nranks):
        rg = rngs[r]
        gi = [np.arange(rg[a][0], rg[a][1]) for a in range(3)]
        lc = [g.size for g in gi]
        own1 = [(gi[a] >= rg[a][2]) & (gi[a] < rg[a][3]) for a in range(3)]
        own = (own1[2][:, None, None] & own1[1][None, :, None] & own1[0][None, None, :]).reshape(-1)
        K, J, I = np.meshgrid(gi[2], gi[1], gi[0], indexing="ij")
        K, J, I = K.reshape(-1)[~own], J.reshape(-1)[~own], I.reshape(-1)[~own]
        oc = [owner_coord[0][I], owner_coord[1][J], owner_coord[2][K]]
        orank = oc[0] + part3[0] * (oc[1] + part3[1] * oc[2])
        sidx = np.zeros(orank.size, dtype=np.int64)
        for q in np.unique(orank):
            m = orank == q
            rq = rngs[int(q)]
            nxq, nyq = rq[0][1] - rq[0][0], rq[1][1] - rq[1][0]
            sidx[m] = (I[m] - rq[0][0]) + nxq * ((J[m] - rq[1][0]) + nyq * (K[m] - rq[2][0]))
        owners.append(np.ascontiguousarray(own, dtype=np.uint8))
        dsts.append(np.ascontiguousarray(np.flatnonzero(~own), dtype=np.int32))
        src_ranks.append(np.ascontiguousarray(orank, dtype=np.int32))
        src_idxs.append(np.ascontiguousarray(sidx, dtype=np.int32))
        ns.append(int(np.prod(lc)))
    rps = [np.ascontiguousarray(s_[0], dtype=np.int32) for s_ in systems]
    cis = [np.ascontiguousarray(s_[1], dtype=np.int32) for s_ in systems]
    vals = [np.ascontiguousarray(s_[2], dtype=np.float64) for s_ in systems]
    rhss = [np.ascontiguousarray(s_[3], dtype=np.float64) for s_ in systems]
    xs = [np.zeros(ns[r] * b) for r in range(nranks)]
    assert all(rps[r].size == ns[r] + 1 for r in range(nranks))

    def ptrs(arrs, ctype):
        return (C.POINTER(ctype) * nranks)(*[a.ctypes.data_as(C.POINTER(ctype)) for a in arrs])

    its, ach, secs = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
    L = O.lib()
    L.orc_schwarz_ilu0_bicgstab.restype = C.c_int
    st = L.orc_schwarz_ilu0_bicgstab(
        C.c_int(nranks), C.c_int(b), (C.c_int * nranks)(*ns), ptrs(rps, C.c_int), ptrs(cis, C.c_int), ptrs(vals, C.c_double),
        ptrs(xs, C.c_double), ptrs(rhss, C.c_double), ptrs(owners, C.c_ubyte), (C.c_int * nranks)(*[d.size for d in dsts]),
        ptrs(dsts, C.c_int), ptrs(src_ranks, C.c_int), ptrs(src_idxs, C.c_int), C.c_double(reduction), C.c_int(maxit),
        C.byref(its), C.byref(ach), C.byref(secs))
    return xs, int(st), its.value, ach.value, secs.value


def gather_owned(results, cells, nranks, b, part=None):
    """Assemble the owned blocks of per-rank local vectors (x fastest) into the global vector."""
    dim = len(cells)
    part = tuple(part) if part is not None else problems.default_partitioning(dim, nranks)
    c3 = tuple(cells) + (1,) * (3 - dim)
    out = np.zeros((c3[2], c3[1], c3[0], b))
    for r, v in enumerate(results):
        rng = problems.box_partition(cells, part, r) + [(0, 1, 0, 1)] * (3 - dim)
        shape = tuple(rng[a][1] - rng[a][0] for a in (2, 1, 0)) + (b,)
        gs = tuple(slice(rng[a][2], rng[a][3]) for a in (2, 1, 0))
        ls = tuple(slice(rng[a][2] - rng[a][0], rng[a][3] - rng[a][0]) for a in (2, 1, 0))
        out[gs] = np.asarray(v).reshape(shape)[ls]
    return out.reshape(-1)
