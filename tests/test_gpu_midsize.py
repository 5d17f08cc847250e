"""Parity of iteration COUNTS and fields at sizes between the unit-test grids and the benchmark grid (nothing
checked BiCGSTAB / Newton counts beyond 48x32 and 20x12x10):

  * 2p lens (BASELINE config 3) at 96^3 -- the size bench.py's CPU leg runs: one Newton iteration, BiCGSTAB count EQUAL to the
    oracle's, update to 1e-8; at 64^3 a full Newton solve: every BiCGSTAB count, the Newton count and the fields;
  * 1p compressible with tabulated water and log-normal K (config 2) at 96^3: one Newton iteration, count and update;
  * tracer transport (config 5) at 128^3 on the velocity field of a 1p solve: volume fluxes bit-identical, explicit and
    implicit steps against the oracle.

The oracle side runs oracle/dist_oracle.single_rank: dune-istl's BiCGSTAB sequence with every scalar product summed in the
device's reduction tree (dist_oracle.gpu_sum), so both sides execute the identical iteration and the counts are compared with ==.
"""
import numpy as np
import pytest

from dumux_b200 import binding as B
from dumux_b200 import problems
from oracle import dist_oracle as D
from oracle.oracle_py import Oracle

pytestmark = pytest.mark.gpu


def _one_newton_iteration(ro, u0, prev, maxit=2000):
    res, jac = ro.o.assemble(u0, prev)
    dx, st, its, red = ro.bicgstab(jac, res, 1e-6, maxit)
    return u0 - dx, st, its


def test_2p_lens_96_newton_iteration_bicgstab_count_equals_oracle(engine_factory):
    spec = problems.twop_lens((96, 96, 96), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
    u0 = spec.initial.reshape(-1).copy()
    uo, sto, its_o = _one_newton_iteration(D.single_rank(spec), u0, u0)
    e = engine_factory(spec)
    e.upload(B.VEC_CUR, u0)
    e.upload(B.VEC_PREV, u0)
    stg, its_g, shift, *_ = e.newton_step(e.newton_params(lin_maxit=2000))
    ug = e.download(B.VEC_CUR)
    assert sto == 0 and stg == 0
    assert its_g == its_o, (its_g, its_o)
    assert its_g > 100                                   # the regime the unit-test grids never reach
    up, us = ug.reshape(-1, 2), uo.reshape(-1, 2)
    assert np.linalg.norm(up[:, 0] - us[:, 0]) <= 1e-8 * np.linalg.norm(us[:, 0])
    assert np.linalg.norm(up[:, 1] - us[:, 1]) <= 1e-8 * max(1.0, np.linalg.norm(us[:, 1]))


def test_2p_lens_64_newton_solve_all_counts_equal_oracle(engine_factory):
    spec = problems.twop_lens((64, 64, 64), law="bc", heterogeneity_sigma=0.5, dt=250.0, plane_rng=True)
    uo, sto, nsteps_o, lin_o = D.single_rank(spec).newton(spec.initial, spec.initial, lin_maxit=2000)
    e = engine_factory(spec)
    ug, stg, rep = e.newton(spec.initial, spec.initial, lin_maxit=2000)
    lin_g = [rep.linear_iterations[i] for i in range(rep.newton_iterations)]
    assert sto == 0 and stg == 0 and rep.newton_iterations == nsteps_o
    assert lin_g == lin_o, (lin_g, lin_o)
    up, us = ug.reshape(-1, 2), uo.reshape(-1, 2)
    assert np.linalg.norm(up[:, 0] - us[:, 0]) <= 1e-8 * np.linalg.norm(us[:, 0])
    assert np.linalg.norm(up[:, 1] - us[:, 1]) <= 1e-8 * max(1.0, np.linalg.norm(us[:, 1]))


def test_1p_compressible_96_newton_iteration_equals_oracle(engine_factory):
    """BASELINE config 2 at 96^3: tabulated IAPWS water, log-normal K (the mt19937 replay of examples/1ptracer), dt 0.002 s"""
    spec = problems.onep_compressible((96, 96, 96), lognormal=True, dt=0.002)
    u0 = spec.initial.reshape(-1).copy()
    uo, sto, its_o = _one_newton_iteration(D.single_rank(spec), u0, u0)
    e = engine_factory(spec)
    e.upload(B.VEC_CUR, u0)
    e.upload(B.VEC_PREV, u0)
    stg, its_g, shift, *_ = e.newton_step(e.newton_params(lin_maxit=2000))
    ug = e.download(B.VEC_CUR)
    assert sto == 0 and stg == 0
    assert its_g == its_o, (its_g, its_o)
    assert np.linalg.norm(ug - uo) <= 1e-8 * np.linalg.norm(uo)


def test_tracer_128_steps_equal_oracle(engine_factory):
    """BASELINE config 5 at 128^3: frozen velocity field from a 1p pressure field, explicit steps (diagonal Jacobian) and one
    implicit step (7-point Jacobian, ILU0-BiCGSTAB) against the oracle."""
    cells = (128, 128, 128)
    n = int(np.prod(cells))
    ps = problems.onep_tracer_pressure((4, 4, 4))
    ps = problems.ProblemSpec(**{**ps.__dict__, "cells": cells})
    ctr = problems.cell_centers(cells, ps.lower, ps.upper)
    lens = problems._in_box(ctr, [0.2] * 3, [0.8] * 3, 1.5e-7)
    ps.K = np.where(lens, 1e-11, 1e-10) * problems.fast_lognormal_multiplier(n, 0.5, 0)
    ps.phi = np.full(n, 0.2)
    ps.region = np.zeros(n, dtype=np.int32)
    ps.initial = np.zeros((n, 1))
    ps.bc_type, ps.bc_values = {}, {}
    for side in range(6):
        fc = problems.side_face_centers(cells, ps.lower, ps.upper, side)
        z = fc[:, 2]
        d = (z < 1e-6) | (z > 1.0 - 1e-6)
        ps.bc_type[side] = np.where(d, problems.BC_DIRICHLET, problems.BC_NEUMANN).astype(np.int32)
        v = np.zeros((fc.shape[0], 1))
        v[d, 0] = 1.0e5 * (1.1 - z[d] * 0.1)
        ps.bc_values[side] = v
    rng = np.random.RandomState(2)
    p = 1.0e5 * (1.1 - 0.1 * ctr[:, 2]) + rng.uniform(-5.0, 5.0, size=n)           # mostly upward flow + noise
    e1 = engine_factory(ps)
    vf_g = e1.volume_flux(p)
    vf_o = Oracle(ps).volume_flux(p)
    assert np.array_equal(vf_g, vf_o)
    for implicit, dt, steps in ((False, 0.01, 3), (True, 5.0, 1)):
        ts = problems.tracer_transport(cells, vf_o, dt=dt, implicit=implicit)
        ts.initial[:, 0] = rng.uniform(0.0, 2e-11, size=n)
        ro = D.single_rank(ts)
        x = ts.initial.reshape(-1).copy()
        its_o = []
        for _ in range(steps):
            r, j = ro.o.assemble(x, x)
            dx, st, its, red = ro.bicgstab(j, r, 1e-10, 500)
            assert st == 0
            its_o.append(its)
            x = x - dx
        et = engine_factory(ts)
        et.upload(B.VEC_CUR, ts.initial)
        et.upload(B.VEC_PREV, ts.initial)
        prm = et.newton_params(lin_reduction=1e-10, lin_maxit=500)
        its_g = []
        for _ in range(steps):
            st, its, *_ = et.newton_step(prm)
            assert st == 0
            its_g.append(its)
            et.advance_timestep()
        xg = et.download(B.VEC_CUR)
        assert its_g == its_o, (implicit, its_g, its_o)
        assert np.abs(xg - x).max() <= 1e-12 * np.abs(x).max()
        assert np.abs(x - ts.initial.reshape(-1)).max() > 1e-7 * np.abs(x).max()   # the field really moved
