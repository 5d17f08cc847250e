"""Oracle restatement of the dune-istl arithmetic checked against independent references:
scipy direct solves, a dense textbook ILU(0), and test/linear/test_linearsolver.cc's set-up (2x2-block Laplacian; the
reference asserts only `converged`, so the numbers here are pinned against scipy instead)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle.oracle_py as O
from dumux_b200 import problems


def _block_laplacian(N, b, seed=0):
    """dune-istl setupLaplacian pattern on an N x N grid with b x b blocks: diag block = 4 I (+ coupling), off = -I."""
    rng = np.random.RandomState(seed)
    n = N * N
    rows, cols, blocks = [], [], []
    for j in range(N):
        for i in range(N):
            I = i + N * j
            nb = []
            if j > 0: nb.append(I - N)
            if i > 0: nb.append(I - 1)
            nb.append(I)
            if i + 1 < N: nb.append(I + 1)
            if j + 1 < N: nb.append(I + N)
            for J in nb:
                rows.append(I); cols.append(J)
                if J == I:
                    blk = 4.0 * np.eye(b) + 0.3 * rng.uniform(-1, 1, (b, b))
                else:
                    blk = -np.eye(b) + 0.1 * rng.uniform(-1, 1, (b, b))
                blocks.append(blk)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    for r in rows:
        rowptr[r + 1] += 1
    rowptr = np.cumsum(rowptr).astype(np.int32)
    colidx = np.array(cols, dtype=np.int32)
    values = np.array(blocks).reshape(-1)
    return n, rowptr, colidx, values


def _to_scipy(n, b, rowptr, colidx, values):
    return sp.bsr_matrix((values.reshape(-1, b, b), colidx, rowptr), shape=(n * b, n * b)).tocsr()


def _dense_ilu0(A, pattern):
    """Textbook IKJ ILU(0) on a dense copy restricted to `pattern` (scalar version)."""
    n = A.shape[0]
    LU = A.copy()
    for i in range(1, n):
        for k in range(i):
            if not pattern[i, k]:
                continue
            LU[i, k] /= LU[k, k]
            for j in range(k + 1, n):
                if pattern[i, j]:
                    LU[i, j] -= LU[i, k] * LU[k, j]
    return LU


@pytest.mark.parametrize("b", [1, 2])
def test_spmv_matches_scipy(b):
    n, rp, ci, v = _block_laplacian(7, b)
    x = np.random.RandomState(1).uniform(-1, 1, n * b)
    y = O.spmv(n, b, rp, ci, v, x)
    assert np.allclose(y, _to_scipy(n, b, rp, ci, v) @ x, rtol=1e-14, atol=1e-14)


def test_ilu0_scalar_matches_textbook():
    n, rp, ci, v = _block_laplacian(6, 1)
    A = _to_scipy(n, 1, rp, ci, v).toarray()
    LU = _dense_ilu0(A, A != 0)
    ilu, st = O.ilu0_factor(n, 1, rp, ci, v)
    assert st == 0
    for i in range(n):
        for k in range(rp[i], rp[i + 1]):
            j = ci[k]
            # dune-istl stores L (unit lower, entries already multiplied by U_jj^-1), U, and the INVERSE diagonal
            expect = 1.0 / LU[i, i] if i == j else LU[i, j]
            assert ilu[k] == pytest.approx(expect, rel=1e-13)
    # apply: v = U^-1 L^-1 d
    d = np.random.RandomState(2).uniform(-1, 1, n)
    Lm = np.tril(LU, -1) + np.eye(n)
    Um = np.triu(LU)
    assert np.allclose(O.ilu0_apply(n, 1, rp, ci, ilu, d), np.linalg.solve(Um, np.linalg.solve(Lm, d)), rtol=1e-12)


def test_ilu0_block_is_exact_for_block_tridiagonal():
    """ILU(0) of a block-tridiagonal matrix has no dropped fill: (LU)^-1 d must solve A v = d exactly."""
    b, n = 2, 12
    rng = np.random.RandomState(3)
    rows = []
    rp = [0]
    ci, blocks = [], []
    for i in range(n):
        for j in (i - 1, i, i + 1):
            if 0 <= j < n:
                ci.append(j)
                blocks.append(5.0 * np.eye(2) + rng.uniform(-1, 1, (2, 2)) if i == j else rng.uniform(-1, 1, (2, 2)))
        rp.append(len(ci))
    rp, ci, v = np.array(rp, dtype=np.int32), np.array(ci, dtype=np.int32), np.array(blocks).reshape(-1)
    ilu, st = O.ilu0_factor(n, b, rp, ci, v)
    d = rng.uniform(-1, 1, n * b)
    x = O.ilu0_apply(n, b, rp, ci, ilu, d)
    assert np.allclose(_to_scipy(n, b, rp, ci, v) @ x, d, rtol=1e-12, atol=1e-12)


def test_ilu0_singular_block_reported():
    n, rp, ci, v = _block_laplacian(3, 2)
    v = v.copy().reshape(-1, 4)
    v[[k for k in range(rp[0], rp[1]) if ci[k] == 0][0]] = 0.0
    ilu, st = O.ilu0_factor(n, 2, rp, ci, v.reshape(-1))
    assert st != 0


@pytest.mark.parametrize("b,N", [(1, 12), (2, 2), (2, 15)])
def test_bicgstab_converges_to_direct_solution(b, N):
    """test_linearsolver.cc uses ProblemSize = 2 with 2x2 blocks; also a larger grid."""
    n, rp, ci, v = _block_laplacian(N, b)
    A = _to_scipy(n, b, rp, ci, v)
    rhs = np.random.RandomState(4).uniform(-1, 1, n * b)
    L = O.lib()
    x = np.zeros(n * b)
    import ctypes as C
    its, red = C.c_int(0), C.c_double(0)
    st = L.orc_ilu0_bicgstab(n, b, rp, ci, v, x, rhs, 1e-13, 250, C.byref(its), C.byref(red))
    assert st == 0 and red.value < 1e-13 and 1 <= its.value < 60
    assert np.allclose(x, spla.spsolve(A.tocsc(), rhs), rtol=1e-10, atol=1e-12)
    # the reported reduction is ||b - A x|| / ||b|| (x0 = 0)
    assert np.linalg.norm(rhs - A @ x) / np.linalg.norm(rhs) == pytest.approx(red.value, rel=1e-3, abs=1e-15)


def test_bicgstab_stopping_rules():
    n, rp, ci, v = _block_laplacian(10, 2)
    rhs = np.random.RandomState(5).uniform(-1, 1, n * 2)
    import ctypes as C
    L = O.lib()
    # maxit reached -> status 1, iterations == maxit
    x = np.zeros(n * 2); its, red = C.c_int(0), C.c_double(0)
    assert L.orc_ilu0_bicgstab(n, 2, rp, ci, v, x, rhs, 1e-13, 2, C.byref(its), C.byref(red)) == 1 and its.value == 2
    # zero right-hand side: converged at once (norm < 1e-30), no iterations, x untouched
    x = np.zeros(n * 2)
    assert L.orc_ilu0_bicgstab(n, 2, rp, ci, v, x, np.zeros(n * 2), 1e-6, 250, C.byref(its), C.byref(red)) == 0
    assert its.value == 0 and not x.any()
    # initial guess is honoured: the reduction is relative to the INITIAL defect b - A x0 (dune-istl semantics), so a
    # start at the solution still iterates on the rounding-level defect, but stays at the solution
    xs = spla.spsolve(_to_scipy(n, 2, rp, ci, v).tocsc(), rhs)
    x = xs.copy()
    assert L.orc_ilu0_bicgstab(n, 2, rp, ci, v, x, rhs, 1e-6, 250, C.byref(its), C.byref(red)) in (0, 2)
    assert np.linalg.norm(x - xs) <= 1e-12 * np.linalg.norm(xs)
    # non-finite input -> status 3
    bad = rhs.copy(); bad[3] = np.nan
    x = np.zeros(n * 2)
    assert L.orc_ilu0_bicgstab(n, 2, rp, ci, v, x, bad, 1e-6, 250, C.byref(its), C.byref(red)) == 3


@pytest.mark.parametrize("b,N,restart", [(1, 12, 10), (2, 2, 10), (2, 15, 10), (2, 15, 3), (2, 15, 40)])
def test_gmres_converges_to_direct_solution(b, N, restart):
    """ILURestartedGMResIstlSolver restatement: converges to the direct solution for every restart length; the reported
    reduction is that of the PRECONDITIONED defect (left preconditioning), monotone in the iteration limit."""
    n, rp, ci, v = _block_laplacian(N, b)
    A = _to_scipy(n, b, rp, ci, v)
    rhs = np.random.RandomState(4).uniform(-1, 1, n * b)
    import ctypes as C
    L = O.lib()
    x = np.zeros(n * b); its, red = C.c_int(0), C.c_double(0)
    st = L.orc_ilu0_gmres(n, b, rp, ci, v, x, rhs, 1e-13, 400, restart, C.byref(its), C.byref(red))
    assert st == 0 and red.value < 1e-13 and 1 <= its.value < 400
    assert np.allclose(x, spla.spsolve(A.tocsc(), rhs), rtol=1e-9, atol=1e-11)
    # preconditioned defect: ||M^-1 (b - A x)|| / ||M^-1 b||
    ilu, _ = O.ilu0_factor(n, b, rp, ci, v)
    pd = lambda r: np.linalg.norm(O.ilu0_apply(n, b, rp, ci, ilu, r))
    assert pd(rhs - A @ x) / pd(rhs) <= 5e-13
    # residual norms are non-increasing in the iteration limit (minimal-residual property inside a cycle, restarts keep x)
    last = 1.0
    for maxit in (1, 2, 4, 8):
        x = np.zeros(n * b)
        stm = L.orc_ilu0_gmres(n, b, rp, ci, v, x, rhs, 1e-13, maxit, restart, C.byref(its), C.byref(red))
        assert stm in (0, 1) and its.value <= maxit and red.value <= last * (1 + 1e-12)
        last = red.value


def test_gmres_full_cycle_matches_scipy_gmres():
    """One un-restarted cycle of left-preconditioned GMRes minimises ||M^-1(b - A x)|| over the Krylov space: compare the iterate
    after k steps with scipy's GMRes on the explicitly preconditioned system."""
    n, rp, ci, v = _block_laplacian(8, 2)
    A = _to_scipy(n, 2, rp, ci, v)
    rhs = np.random.RandomState(9).uniform(-1, 1, n * 2)
    ilu, _ = O.ilu0_factor(n, 2, rp, ci, v)
    Minv = np.column_stack([O.ilu0_apply(n, 2, rp, ci, ilu, e) for e in np.eye(n * 2)])
    import ctypes as C
    x = np.zeros(n * 2); its, red = C.c_int(0), C.c_double(0)
    k = 6
    O.lib().orc_ilu0_gmres(n, 2, rp, ci, v, x, rhs, 1e-30, k, 50, C.byref(its), C.byref(red))
    # minimiser over span{M^-1 b, (M^-1 A) M^-1 b, ...} by dense least squares
    B = Minv @ A.toarray()
    c = Minv @ rhs
    Kry = [c]
    for _ in range(k - 1):
        Kry.append(B @ Kry[-1])
    Q, _ = np.linalg.qr(np.column_stack(Kry))
    y, *_ = np.linalg.lstsq(B @ Q, c, rcond=None)
    assert np.linalg.norm(x - Q @ y) <= 1e-8 * np.linalg.norm(x)


@pytest.mark.parametrize("b", [1, 2])
def test_ssor_is_the_symmetric_gauss_seidel_splitting(b):
    """SeqSSOR(1, w = 1) from v = 0 equals (D+U)^-1 D (D+L)^-1 d with BLOCK diagonal D (dune-istl bsorf/bsorb)."""
    n, rp, ci, v = _block_laplacian(7, b)
    A = _to_scipy(n, b, rp, ci, v).toarray()
    N = n * b
    blk = np.arange(N) // b
    D = np.where(blk[:, None] == blk[None, :], A, 0.0)
    Lo = np.where(blk[:, None] > blk[None, :], A, 0.0)
    Up = np.where(blk[:, None] < blk[None, :], A, 0.0)
    d = np.random.RandomState(3).standard_normal(N)
    ref = np.linalg.solve(D + Up, D @ np.linalg.solve(D + Lo, d))
    assert np.allclose(O.ssor_apply(n, b, rp, ci, v, d), ref, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("krylov", ["cg", "bicgstab"])
@pytest.mark.parametrize("b,N", [(1, 12), (2, 15)])
def test_ssor_krylov_converges_to_direct_solution(krylov, b, N):
    """SSORCGIstlSolver / SSORBiCGSTABIstlSolver restatements on the SPD block Laplacian of test_linearsolver.cc."""
    n, rp, ci, v = _block_laplacian(N, b)
    if krylov == "cg":
        # CG needs a symmetric positive definite operator: symmetrise the randomly perturbed blocks (the pattern is symmetric)
        Ad = _to_scipy(n, b, rp, ci, v).toarray()
        Ad = 0.5 * (Ad + Ad.T)
        vb = np.empty((len(ci), b, b))
        for i in range(n):
            for k in range(rp[i], rp[i + 1]):
                vb[k] = Ad[i * b:(i + 1) * b, ci[k] * b:(ci[k] + 1) * b]
        v = vb.reshape(-1)
    A = _to_scipy(n, b, rp, ci, v)
    rhs = np.random.RandomState(4).uniform(-1, 1, n * b)
    x, st, its, red = O.ssor_solve(n, b, rp, ci, v, rhs, krylov, 1e-13, 400)
    assert st == 0 and red < 1e-13 and 1 <= its < 200
    assert np.allclose(x, spla.spsolve(A.tocsc(), rhs), rtol=1e-9, atol=1e-11)
    assert np.linalg.norm(rhs - A @ x) / np.linalg.norm(rhs) == pytest.approx(red, rel=1e-3, abs=1e-15)
    # CG: the energy-norm error decreases monotonically with the iteration limit
    if krylov == "cg":
        xs = spla.spsolve(A.tocsc(), rhs)
        last = np.inf
        for maxit in (1, 2, 4, 8, 16):
            xm, stm, itm, _ = O.ssor_solve(n, b, rp, ci, v, rhs, "cg", 1e-30, maxit)
            e = xm - xs
            en = float(e @ (A @ e))
            assert itm == maxit and en <= last * (1 + 1e-12)
            last = en


def test_assembled_jacobian_solve_against_scipy():
    """The real thing: 2p lens Jacobian + residual, ILU0-BiCGSTAB at Newton's reduction vs a sparse direct solve."""
    spec = problems.twop_lens((24, 16), law="vg")
    o = O.Oracle(spec)
    rng = np.random.RandomState(6)
    cur = spec.initial.copy()
    cur[:, 1] = rng.uniform(0, 0.2, cur.shape[0])
    res, jac = o.assemble(cur, spec.initial)
    x, st, its, red = o.solve(jac, res, reduction=1e-12)
    assert st == 0
    xd = spla.spsolve(_to_scipy(o.n, 2, o.rowptr, o.colidx, jac).tocsc(), res)
    assert np.linalg.norm(x - xd) <= 1e-8 * np.linalg.norm(xd)


def test_max_relative_shift_and_norms():
    """newtonsolver.hh:111-129: shift = max |u1-u2| / max(1, |u1+u2|/2)."""
    u1 = np.array([1e5, 0.2, 2e5, 0.0])
    u2 = np.array([1e5 + 1.0, 0.25, 2e5, 1e-9])
    L = O.lib()
    assert L.orc_max_relative_shift(4, u1, u2) == pytest.approx(0.05)
    a = np.arange(1.0, 6.0)
    assert L.orc_norm2(5, a) == pytest.approx(np.sqrt(55.0)) and L.orc_dot(5, a, a) == 55.0


# test/linear/test_parallel_amg_smoothers.cc restated on the oracle: AMG with Dumux::ParMTSSOR / ParMTSOR / ParMTJac smoothers
# (2 iterations, relaxation 0.8) preconditioning CG on the CCTpfa Helmholtz operator of a 300 x 300 grid, b = A 1, reduction 1e-15,
# at most 200 iterations; the reference accepts "converged and | |x|^2 - N | <= 1e-10 N"
@pytest.mark.parametrize("smoother", ["par_mt_ssor", "par_mt_sor", "par_mt_jac"])
def test_parallel_amg_smoothers_helmholtz_oracle(smoother):
    from dumux_b200 import problems
    from oracle import dist_oracle as D
    N = 300
    spec = problems.onep_incompressible((N, N))
    ro = D.single_rank(spec, gpu_reduction=False)
    rp, ci = ro.o.rowptr, ro.o.colidx
    n, h = N * N, 1.0 / N
    rows = np.repeat(np.arange(n), np.diff(rp))
    vals = np.where(ci == rows, 0.0, -1.0)
    vals[ci == rows] = (np.diff(rp) - 1) + h * h
    b = O.spmv(n, 1, rp, ci, vals, np.ones(n))
    ro.amg_params = dict(smoother=smoother, smoother_iterations=2, smoother_relaxation=0.8)
    x, st, its, red = ro.cg(vals, b, 1e-15, 200, precond="amg")
    assert st == 0 and its < 60
    assert abs(np.dot(x, x) - n) <= 1e-10 * n
