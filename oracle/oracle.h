/*
 * oracle/oracle.h -- C API of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The oracle is a CPU restatement of DuMux's Newton-step hot path (cell-centred TPFA residual and
 * numerically differentiated Jacobian + ILU0/BiCGSTAB + Newton control), written from the reference
 * sources cited function-by-function in oracle.cpp.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product library
 * (dumux_b200/csrc) never links, includes or calls anything in this directory.
 *
 * Parity status: the reference itself (DuMux + DUNE 2.10 + MPI) cannot be built in this image, so
 * the oracle is pinned against the reference's golden VTU fields (tests/golden/, extracted by
 * tests/golden/make_golden.py) at the reference's own fuzzy tolerance (rel 1e-2 / abs 1.5e-7 on
 * Float32 data) and against closed-form known answers; the dune-istl arithmetic (ILU0, BiCGSTAB) is
 * restated from the published algorithm (DUNE 2.10) and is "parity unpinned" at tight tolerance.
 */
#ifndef ORACLE_H
#define ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_problem orc_problem;

typedef struct {
    int    enable_gravity;       /* Problem.EnableGravity, default 1 (dumux/common/parameters.cc:243) */
    double gravity;              /* 9.81, acts along -e_{dim-1} (dumux/common/fvspatialparams.hh:48-52) */
    double upwind_weight;        /* Flux.UpwindWeight, default 1.0 (parameters.cc:236) */
    int    fd_method;            /* Assembly.NumericDifferenceMethod: 1 fwd (default), 0 central, -1 bwd, 5 five-point */
    double base_eps;             /* Assembly.NumericDifference.BaseEpsilon, default 1e-10 */
    double privar_magnitude[2];  /* Assembly.NumericDifference.PriVarMagnitude, <=0: unset */
    int    stationary;           /* 1: no storage term */
    double dt;                   /* time-step size for the storage term */
    double extrusion;            /* constant extrusion factor */
    int    use_std_pow;          /* oracle-only switch: 1 -> glibc pow instead of the deterministic pow */
    int    num_threads;          /* OpenMP threads for assembly (<=0: all) */
} orc_options;

enum { ORC_MODEL_1P = 1, ORC_MODEL_2P = 2 };
enum { ORC_LAW_BROOKSCOREY = 0, ORC_LAW_VANGENUCHTEN = 1 };
/* fd_method value selecting DiffMethod::analytic (incompressible 1p only: 1p/incompressiblelocalresidual.hh) */
enum { ORC_DIFF_ANALYTIC = 100 };
enum { ORC_SOLVER_BICGSTAB = 0, ORC_SOLVER_GMRES = 1 };   /* ILUBiCGSTABIstlSolver / ILURestartedGMResIstlSolver */
enum { ORC_BC_NEUMANN = 0, ORC_BC_DIRICHLET = 1, ORC_BC_NONE = 2, ORC_BC_OUTFLOW = 3 };
/* sides: 0 -x, 1 +x, 2 -y, 3 +y, 4 -z, 5 +z (YaspGrid indexInInside) */

void orc_default_options(orc_options* o);
orc_problem* orc_create(int model, int dim, const int* cells, const double* lower, const double* upper);
/* tensor-product variant: explicit node coordinates per axis (cells[a]+1 values each) */
orc_problem* orc_create_tensor(int model, int dim, const int* cells, const double* x, const double* y, const double* z);
void orc_destroy(orc_problem* p);
int  orc_num_cells(const orc_problem* p);
int  orc_num_eq(const orc_problem* p);
void orc_set_options(orc_problem* p, const orc_options* o);
void orc_set_cell_fields(orc_problem* p, const double* K, const double* phi, const int* region);
void orc_set_permeability_diagonal(orc_problem* p, const double* kx, const double* ky, const double* kz);
void orc_set_source(orc_problem* p, const double* q);
/* BC: params = {pcEntry, lambda}, reg = {pcLowSwe}; VG: params = {alpha, n, l}, reg = {pcLowSwe, pcHighSwe, krnLowSwe, krwHighSwe} */
void orc_set_material(orc_problem* p, int region, int law, const double* params, double swr, double snr,
                      int regularize, const double* reg);
/* FVSpatialParams::wettingPhase per region (2p/volumevariables.hh:132; 0 = phase 0 wets, the default) */
void orc_set_wetting_phase(orc_problem* p, int region, int phase);
void orc_set_fluids(orc_problem* p, const double* rho, const double* mu);
/* tabulated liquid (dumux/material/components/tabulatedcomponent.hh): tables values[iT + iP*nT], per-temperature pressure range pmin[nT], pmax[nT] */
void orc_set_fluid_table(orc_problem* p, int nT, int nP, double Tmin, double Tmax, const double* pmin, const double* pmax,
                         const double* rho, const double* mu, double temperature);
/* number of faces on a side, and the per-side arrays: type[nf], values[nf*numEq] (Dirichlet priVars or Neumann fluxes) */
int  orc_side_faces(const orc_problem* p, int side);
void orc_set_boundary(orc_problem* p, int side, const int* type, const double* values);
/* centre of boundary face f of a side (for building BC arrays): out[3] */
void orc_side_face_center(const orc_problem* p, int side, int f, double* out);
void orc_cell_center(const orc_problem* p, int cell, double* out);

int  orc_pattern_nnz(const orc_problem* p);
void orc_pattern(const orc_problem* p, int* rowptr, int* colidx);
/* residual[n*b], jac[nnzb*b*b] (either may be NULL); prev may be NULL if stationary */
void orc_assemble(orc_problem* p, const double* cur, const double* prev, double* residual, double* jac);
/* secondary variables for output/golden comparison: out[n*12] = Sw,Sn,pw,pn,rhow,rhon,mobw,mobn,pc,porosity,K,0 */
void orc_volvars(orc_problem* p, const double* cur, double* out);

/* dune-istl restatement. x: in initial guess / out solution; b is NOT modified. status 0 ok, 1 not converged, 2 breakdown, 3 non-finite */
/* the same solver on `nranks` overlapping-Schwarz ranks (OpenMP threads of this process), see oracle.cpp */
int  orc_schwarz_ilu0_bicgstab(int nranks, int b, const int* n, const int* const* rowptr, const int* const* colidx,
                               const double* const* values, double* const* x, const double* const* rhs,
                               const unsigned char* const* owner, const int* ncopy, const int* const* copy_dst,
                               const int* const* copy_src_rank, const int* const* copy_src_idx, double reduction, int maxit,
                               int* iterations, double* achieved, double* seconds);
int  orc_ilu0_bicgstab(int n, int b, const int* rowptr, const int* colidx, const double* values,
                       double* x, const double* rhs, double reduction, int maxit,
                       int* iterations, double* achieved_reduction);
/* ILURestartedGMResIstlSolver (istlsolvers.hh:660-667): left-preconditioned GMRes(restart); `achieved` refers to the
   PRECONDITIONED defect, which is what Dune::RestartedGMResSolver monitors */
int  orc_ilu0_gmres(int n, int b, const int* rowptr, const int* colidx, const double* values, double* x, const double* rhs,
                    double reduction, int maxit, int restart, int* iterations, double* achieved_reduction);
/* SSORCGIstlSolver (krylov 0) / SSORBiCGSTABIstlSolver (krylov 1), istlsolvers.hh:686-714: SeqSSOR(1, w = 1) preconditioner */
int  orc_ssor_solve(int n, int b, const int* rowptr, const int* colidx, const double* values, double* x, const double* rhs, int krylov,
                    double reduction, int maxit, int* iterations, double* achieved_reduction);
/* v = SeqSSOR(A)(d) from v = 0 (one forward + one backward block Gauss-Seidel sweep) */
void orc_ssor_apply(int n, int b, const int* rowptr, const int* colidx, const double* values, double* v, const double* d);
/* Dumux::ParMTJac (kind 0) / ParMTSOR (1) / ParMTSSOR (2), dumux/linear/preconditioners.hh:330-620 */
int  orc_parmt_solve(int kind, int iterations, double relaxation, int n, int b, const int* rowptr, const int* colidx, const double* values,
                     double* x, const double* rhs, int krylov, double reduction, int maxit, int* its, double* achieved);
void orc_parmt_apply(int kind, int iterations, double relaxation, int n, int b, const int* rowptr, const int* colidx, const double* values,
                     double* v, const double* d);
int  orc_parmt_colors(int n, const int* rowptr, const int* colidx, int* colors);
/* aggregation AMG on the structured hierarchy (oracle/amg_oracle.py restates dumux_b200/csrc/amg.cu) */
void orc_amg_galerkin(int b, int dim, const int* fcells, const int* fown_lo, const int* fown_hi, const int* f_rowptr, const double* fA,
                      const int* ccells, const int* cown_lo, const int* c_rowptr, double* cA);
int  orc_ssor_factor(int n, int b, const int* rowptr, const int* colidx, const double* A, double* out);
/* tracer: binary diffusion coefficient D (FluidSystem::binaryDiffusionCoefficient) and SpatialParams.Tortuosity (default 0.5) of
   DiffusivityConstantTortuosity; D = 0 (the default) switches Fick's law off */
void orc_set_tracer_dispersion(orc_problem* p, const double* disp);     /* n.D.n of the dispersion tensor per cell and side, NULL = off */
void orc_set_tracer_diffusion(orc_problem* p, double D, double tortuosity);
/* linear solver used by orc_newton_solve(_ex) / orc_run_timeloop: ORC_SOLVER_*; restart <= 0: 10 (LinearSolver.GMResRestart) */
void orc_set_linear_solver(orc_problem* p, int kind, int restart);
/* standalone pieces for kernel-level parity */
int  orc_ilu0_factor(int n, int b, const int* rowptr, const int* colidx, const double* values, double* ilu);
void orc_ilu0_apply(int n, int b, const int* rowptr, const int* colidx, const double* ilu, double* v, const double* d);
void orc_spmv(int n, int b, const int* rowptr, const int* colidx, const double* values, const double* x, double* y);
double orc_norm2(int n, const double* v);
double orc_dot(int n, const double* a, const double* b);
double orc_max_relative_shift(int n, const double* u1, const double* u2);

typedef struct {
    int    newton_iterations;
    int    converged;
    int    linear_iterations_total;
    double last_shift;
    double t_assemble, t_solve, t_update;   /* seconds, same three buckets as newtonsolver.hh:950-955 */
    int    linear_iterations[64];
    double shifts[64];
} orc_newton_report;

/* update strategy and convergence criteria of the Newton solver (newtonsolver.hh:1213-1232) */
typedef struct {
    int    use_line_search;               /* Newton.UseLineSearch */
    double line_search_min_relaxation;    /* Newton.LineSearchMinRelaxationFactor 0.125 */
    int    enable_shift_criterion;        /* 1 */
    int    enable_residual_criterion;     /* 0 */
    int    enable_absolute_residual_criterion;
    int    satisfy_residual_and_shift;
    double residual_reduction;            /* Newton.ResidualReduction 1e-5 */
    double max_absolute_residual;         /* Newton.MaxAbsoluteResidual 1e-5 */
} orc_newton_options;
void orc_default_newton_options(orc_newton_options* o);
/* as orc_newton_solve with lineSearchUpdate_ (:1154-1178), computeResidualReduction_ (:869-881), newtonConverged (:657-701);
   relaxation[64] receives the accepted lambda per iteration, reduction the last residual reduction */
int  orc_newton_solve_ex(orc_problem* p, double* u, const double* prev, double lin_reduction, int lin_maxit,
                         double max_rel_shift, int min_steps, int max_steps, const orc_newton_options* opt,
                         orc_newton_report* rep, double* relaxation, double* reduction_out);

/* One Newton solve at fixed dt (newtonsolver.hh:976-1072); u in/out, prev = previous time level */
int  orc_newton_solve(orc_problem* p, double* u, const double* prev, double lin_reduction, int lin_maxit,
                      double max_rel_shift, int min_steps, int max_steps, orc_newton_report* rep);
/* Full instationary run as in test/porousmediumflow/2p/incompressible/main.cc:126-163 (plain TimeLoop + dt control).
   returns number of time steps; newton_its[] receives the Newton count per step (up to max_steps_out). */
int  orc_run_timeloop(orc_problem* p, double* u, double t_end, double dt_initial, double max_dt,
                      int* newton_its, double* dts, int max_steps_out);

/* tracer transport on a frozen velocity field (examples/1ptracer): volume fluxes vf[n*2*dim] from a 1p pressure field
   (main.cc:162-199), and the TracerLocalResidual assembled explicitly (implicit = 0, the example) or implicitly */
void orc_volume_flux(orc_problem* p, const double* pressure, double* vf);
void orc_tracer_assemble(orc_problem* p, const double* vf, const double* cur, const double* prev, int implicit, double rho,
                         double* residual, double* jac);

/* material-law probes for unit tests: which = 0 pc, 1 krw, 2 krn, 3 dpc_dsw, 4 dkrw_dsw, 5 dkrn_dsw */
double orc_law_eval(orc_problem* p, int region, int which, double sw);
double orc_pow(double x, double y);

#ifdef __cplusplus
}
#endif
#endif
