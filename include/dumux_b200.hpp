// dumux_b200.hpp -- header-only C++ host layer over the C ABI (include/dumux_b200.h).
//
// Mirrors, with the same member names, argument meaning and error behaviour, the DuMux interfaces on the Newton-step
// path, so that a DuMux main (and DuMux's own NewtonSolver, which is duck-typed: test/nonlinear/newton/test_newton.cc:31-78)
// can use them in place of the CPU classes:
//
//   dumux_b200::GpuFVAssembler        <->  Dumux::FVAssembler<TypeTag, DiffMethod::numeric>   dumux/assembly/fvassembler.hh:115-390
//   dumux_b200::GpuILUBiCGSTABSolver  <->  Dumux::ILUBiCGSTABIstlSolver<LSTraits, LATraits>   dumux/linear/istlsolvers.hh:202-390,636-642
//   dumux_b200::GpuNewtonSolver       <->  Dumux::NewtonSolver<Assembler, LinearSolver>       dumux/nonlinear/newtonsolver.hh:309-355,976-1072
//   dumux_b200::NumericalProblem      <->  Dumux::NumericalProblem                            dumux/common/exceptions.hh
//
// DUNE is not a dependency: vectors and matrices are the flat layouts Dune::BlockVector / Dune::BCRSMatrix store
// (BlockVector<FieldVector<double,b>> is a contiguous double[n*b]; BCRS blocks row-major, columns ascending).
// INTEGRATION.md shows the adapter that binds the DUNE containers to these classes inside a DuMux application.
// There is no CPU fallback: every call runs CUDA kernels through libdumux_b200.so or throws.
#ifndef DUMUX_B200_HPP
#define DUMUX_B200_HPP

#include <array>
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "dumux_b200.h"

namespace dumux_b200 {

//! recoverable numerical failure: NewtonSolver catches it and halves the time step (newtonsolver.hh:318-354,514-523)
struct NumericalProblem : std::runtime_error {
    using std::runtime_error::runtime_error;
};
//! usage / CUDA / NCCL error (Dune::InvalidStateException and friends)
struct InvalidState : std::runtime_error {
    using std::runtime_error::runtime_error;
};

//! one GPU context (one per rank); shared by assembler, linear solver and Newton solver
class Context {
public:
    explicit Context(int device = 0)
    {
        if (dmx_create(&ctx_, device) != 0 || !ctx_) throw InvalidState("dmx_create failed: no CUDA device (there is no CPU fallback)");
    }
    //! slab-decomposed run: rank/nranks and the ncclUniqueId of rank 0 (replaces gridView.comm(), istlsolvers.hh:192)
    Context(int device, const void* ncclUniqueId, int rank, int nranks)
    {
        if (dmx_create_distributed(&ctx_, device, ncclUniqueId, rank, nranks) != 0 || !ctx_) throw InvalidState("dmx_create_distributed failed");
    }
    ~Context() { dmx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    dmx_ctx* get() const { return ctx_; }
    //! status > 0: numerical condition (-> NumericalProblem unless allowed), < 0: error
    int check(int rc, bool allowStatus = false) const
    {
        if (rc < 0) throw InvalidState(std::string("libdumux_b200: ") + dmx_last_error(ctx_));
        if (rc > 0 && !allowStatus) throw NumericalProblem(std::string("libdumux_b200: ") + dmx_last_error(ctx_));
        return rc;
    }

private:
    dmx_ctx* ctx_ = nullptr;
};

//! Dune::BlockVector<Dune::FieldVector<double,b>> look-alike: contiguous double[n*b]
class BlockVector {
public:
    BlockVector() = default;
    BlockVector(std::size_t n, int b, double value = 0.0) : b_(b), data_(n * b, value) {}
    std::size_t size() const { return b_ ? data_.size() / b_ : 0; }
    int blockSize() const { return b_; }
    double* operator[](std::size_t i) { return data_.data() + i * b_; }
    const double* operator[](std::size_t i) const { return data_.data() + i * b_; }
    double* data() { return data_.data(); }
    const double* data() const { return data_.data(); }
    BlockVector& operator=(double v) { for (auto& x : data_) x = v; return *this; }
    BlockVector& operator-=(const BlockVector& o) { for (std::size_t i = 0; i < data_.size(); ++i) data_[i] -= o.data_[i]; return *this; }
    double two_norm() const { double s = 0; for (double x : data_) s += x * x; return std::sqrt(s); }

private:
    int b_ = 0;
    std::vector<double> data_;
};

//! Dune::BCRSMatrix<Dune::FieldMatrix<double,b,b>> look-alike
struct BCRSMatrix {
    int n = 0, b = 0;
    std::vector<int> rowptr, colidx;
    std::vector<double> values;          // nnzb*b*b, row-major blocks
    std::size_t nonzeroes() const { return colidx.size(); }
};

//! stands for Dumux::PartialReassembler<Assembler> (assembly/partialreassembler.hh): partial reassembly is not implemented on
//! the device (parity runs set Newton.EnablePartialReassembly = false), so the only value accepted is nullptr
struct PartialReassembler;

//! what GridGeometry + Problem + SpatialParams provide, sampled into flat arrays (common/fvproblem.hh:126-283,
//! porousmediumflow/fvspatialparams.hh:83-99); cells numbered x fastest, boundary faces per side lower axis fastest
struct ProblemData {
    int model = DMX_MODEL_1P, dim = 2;
    std::array<int, 3> cells{{1, 1, 1}};
    std::array<double, 3> lower{{0, 0, 0}}, upper{{1, 1, 1}};
    std::vector<double> permeability, porosity;      // per cell (empty: 1e-10 / 0.4)
    std::vector<int> region;                         // per cell material-law index
    struct Material { int law; std::vector<double> params; double swr, snr; bool regularize; std::vector<double> reg; };
    std::vector<Material> materials;
    std::array<double, 2> density{{1000.0, 1460.0}}, viscosity{{1e-3, 5.7e-4}};
    struct Side { std::vector<int> type; std::vector<double> values; };   // values: [face][numEq] Dirichlet priVars or Neumann fluxes
    std::array<Side, 6> boundary;                    // sides -x,+x,-y,+y,-z,+z; empty = no-flow Neumann
    std::vector<double> source;                      // per cell and equation (empty: none)
    //! tabulated liquid of the compressible 1p model: the tables TabulatedComponent<H2O>::init filled
    //! (material/components/tabulatedcomponent.hh:330-345), values[iT + iP*nT]; empty: constant density / viscosity
    struct FluidTable { int nT = 0, nP = 0; double Tmin = 0, Tmax = 0, temperature = 293.15; std::vector<double> pmin, pmax, density, viscosity; };
    FluidTable fluidTable;
    //! tracer model (DMX_MODEL_TRACER): frozen volume fluxes [cell][side] (spatialparams_tracer.hh:91-103) and the time
    //! discretisation of FVAssembler<TracerTypeTag, DiffMethod::analytic, implicit> (examples/1ptracer/main.cc:236)
    std::vector<double> volumeFlux;
    bool tracerImplicit = false;
    //! The adapter that samples a DuMux Problem sets these when the problem overrides the solution-dependent interfaces
    //! neumann(element, fvGeometry, elemVolVars, elemFluxVarsCache, scvf) / source(element, fvGeometry, elemVolVars, scv)
    //! (common/fvproblem.hh:262-283,307-331) instead of neumannAtPos / sourceAtPos, or returns a tensor from
    //! SpatialParams::permeability (porousmediumflow/fvspatialparams.hh:83-99) with OFF-DIAGONAL entries: the kernels take sampled
    //! arrays and a scalar or diagonal K, so the assembler refuses such a problem instead of silently freezing the values
    //! (DESIGN.md "Out of scope").  A diagonal tensor goes into permeabilityDiagonal.
    bool solutionDependentNeumann = false, solutionDependentSource = false, tensorPermeability = false;
    //! K = diag(K_xx, K_yy, K_zz) per cell: one array per grid axis (dmx_set_permeability_diagonal); empty = scalar `permeability`
    std::vector<double> permeabilityDiagonal[3];
    dmx_options options;
    ProblemData() { dmx_default_options(&options); }
};

struct IstlSolverResult {            // linear/istlsolvers.hh:153-163 (Dune::InverseOperatorResult + operator bool)
    int iterations = 0;
    double reduction = 0.0;
    bool converged = false;
    explicit operator bool() const { return converged; }
};

// =====================================================================================================================
class GpuFVAssembler {
public:
    using Scalar = double;
    using JacobianMatrix = BCRSMatrix;
    using SolutionVector = BlockVector;
    using ResidualType = BlockVector;
    using Variables = BlockVector;          // assemblers that do not export grid variables: Variables = SolutionVector (newtonsolver.hh:199)

    //! stationary problems (fvassembler.hh:131)
    GpuFVAssembler(std::shared_ptr<Context> ctx, const ProblemData& problem) : ctx_(std::move(ctx)), stationary_(true) { init_(problem); }
    //! instationary problems (fvassembler.hh:154): time-step size instead of the TimeLoop pointer, observing pointer to prevSol
    GpuFVAssembler(std::shared_ptr<Context> ctx, const ProblemData& problem, double dt, const SolutionVector& prevSol)
    : ctx_(std::move(ctx)), stationary_(false), prevSol_(&prevSol)
    {
        init_(problem);
        setTimeStepSize(dt);
    }

    //! fvassembler.hh:179-207: throws NumericalProblem if the residual is not finite (all ranks agree, :504-509).
    //! `partialReassembler` must be null (see PartialReassembler above).
    void assembleJacobianAndResidual(const SolutionVector& curSol, const PartialReassembler* partialReassembler = nullptr)
    {
        if (partialReassembler) throw InvalidState("GpuFVAssembler: partial reassembly is not supported (Newton.EnablePartialReassembly must be false)");
        upload_(curSol);
        ctx_->check(dmx_assemble(ctx_->get(), 1));
        hostJacobianValid_ = hostResidualValid_ = false;
    }
    void assembleJacobian(const SolutionVector& curSol) { assembleJacobianAndResidual(curSol); }     // :212
    void assembleResidual(const SolutionVector& curSol)                                                // :240
    {
        upload_(curSol);
        ctx_->check(dmx_assemble(ctx_->get(), 0));
        hostResidualValid_ = false;
    }
    void assembleResidual(ResidualType& r, const SolutionVector& curSol)                               // :247
    {
        assembleResidual(curSol);
        r = residual();
    }
    //! the pattern is fixed by the grid (jacobianpattern.hh:27-52); allocates the host mirrors (fvassembler.hh:294)
    void setLinearSystem()
    {
        jac_.n = numDofs(); jac_.b = numEq_;
        jac_.rowptr.resize(jac_.n + 1); jac_.colidx.resize(nnzb_);
        ctx_->check(dmx_pattern(ctx_->get(), jac_.rowptr.data(), jac_.colidx.data()));
        jac_.values.assign(nnzb_ * numEq_ * numEq_, 0.0);
        res_ = ResidualType(numDofs(), numEq_);
    }
    //! host views; the data stays on the device until somebody asks for it (SURVEY 8b "Ownership")
    JacobianMatrix& jacobian()
    {
        if (jac_.rowptr.empty()) setLinearSystem();
        if (!hostJacobianValid_) { ctx_->check(dmx_jacobian_download(ctx_->get(), jac_.values.data())); hostJacobianValid_ = true; }
        return jac_;
    }
    ResidualType& residual()
    {
        if (res_.size() == 0) setLinearSystem();
        if (!hostResidualValid_) { ctx_->check(dmx_vec_download(ctx_->get(), DMX_VEC_RESIDUAL, res_.data())); hostResidualValid_ = true; }
        return res_;
    }
    std::size_t numDofs() const { return static_cast<std::size_t>(dmx_num_cells(ctx_->get())); }
    int numEq() const { return numEq_; }
    //! 1p assembler: volume fluxes over all scvfs of all elements from the pressure field `p`, [element][indexInInside]
    //! -- the element loop of examples/1ptracer/main.cc:162-199, input of TracerTestSpatialParams::setVolumeFlux
    std::vector<double> volumeFlux(const SolutionVector& p, int dim)
    {
        upload_(p);
        std::vector<double> vf(numDofs() * 2 * static_cast<std::size_t>(dim));
        ctx_->check(dmx_volume_flux(ctx_->get(), vf.data()));
        return vf;
    }
    const SolutionVector& prevSol() const { return *prevSol_; }
    void setPreviousSolution(const SolutionVector& u) { prevSol_ = &u; prevUploaded_ = false; }
    //! stands for setTimeLoop / timeLoop->timeStepSize() (fvassembler.hh:347-368)
    void setTimeStepSize(double dt)
    {
        options_.dt = dt;
        ctx_->check(dmx_set_options(ctx_->get(), &options_));
    }
    bool isStationaryProblem() const { return stationary_; }
    void updateGridVariables(const SolutionVector&) {}                    // caches are device-internal (:379)
    void resetTimeStep(const SolutionVector&) { prevUploaded_ = false; }  // :387
    //! after a successful time step: xOld = x; gridVariables->advanceTimeStep()
    void advanceTimeStep() { prevUploaded_ = false; }
    const std::shared_ptr<Context>& context() const { return ctx_; }

private:
    void init_(const ProblemData& p)
    {
        if (p.solutionDependentNeumann || p.solutionDependentSource)
            throw InvalidState("GpuFVAssembler: solution-dependent Neumann fluxes / sources are not supported (only the tracer outflow, DMX_BC_OUTFLOW)");
        if (p.tensorPermeability) throw InvalidState("GpuFVAssembler: permeability tensors with off-diagonal entries are not supported (scalar or diagonal K)");
        dmx_ctx* c = ctx_->get();
        ctx_->check(dmx_grid_structured(c, p.model, p.dim, p.cells.data(), p.lower.data(), p.upper.data()));
        numEq_ = dmx_num_eq(c);
        nnzb_ = static_cast<std::size_t>(dmx_nnz_blocks(c));
        options_ = p.options;
        options_.stationary = stationary_ ? 1 : 0;
        ctx_->check(dmx_set_options(c, &options_));
        ctx_->check(dmx_set_cell_fields(c, p.permeability.empty() ? nullptr : p.permeability.data(),
                                        p.porosity.empty() ? nullptr : p.porosity.data(), p.region.empty() ? nullptr : p.region.data()));
        if (!p.permeabilityDiagonal[0].empty())
            ctx_->check(dmx_set_permeability_diagonal(c, p.permeabilityDiagonal[0].data(),
                                                      p.permeabilityDiagonal[1].empty() ? nullptr : p.permeabilityDiagonal[1].data(),
                                                      p.permeabilityDiagonal[2].empty() ? nullptr : p.permeabilityDiagonal[2].data()));
        for (std::size_t r = 0; r < p.materials.size(); ++r) {
            const auto& m = p.materials[r];
            ctx_->check(dmx_set_material(c, static_cast<int>(r), m.law, m.params.data(), m.swr, m.snr, m.regularize ? 1 : 0,
                                         m.reg.empty() ? nullptr : m.reg.data()));
        }
        ctx_->check(dmx_set_fluids(c, p.density.data(), p.viscosity.data()));
        if (p.fluidTable.nT > 0) {
            const auto& t = p.fluidTable;
            ctx_->check(dmx_set_fluid_table(c, t.nT, t.nP, t.Tmin, t.Tmax, t.pmin.data(), t.pmax.data(), t.density.data(), t.viscosity.data(),
                                            t.temperature));
        }
        if (p.model == DMX_MODEL_TRACER) {
            if (p.volumeFlux.size() != numDofs() * 2 * static_cast<std::size_t>(p.dim)) throw InvalidState("tracer: volumeFlux must hold 2*dim values per cell");
            ctx_->check(dmx_set_volume_flux(c, p.volumeFlux.data()));
            ctx_->check(dmx_set_tracer(c, p.tracerImplicit ? 1 : 0));
        }
        for (int s = 0; s < 2 * p.dim; ++s)
            if (!p.boundary[s].type.empty()) ctx_->check(dmx_set_boundary(c, s, p.boundary[s].type.data(), p.boundary[s].values.data()));
        if (!p.source.empty()) ctx_->check(dmx_set_source(c, p.source.data()));
    }
    void upload_(const SolutionVector& curSol)
    {
        if (curSol.size() != numDofs() || curSol.blockSize() != numEq_) throw InvalidState("solution vector size mismatch");
        ctx_->check(dmx_vec_upload(ctx_->get(), DMX_VEC_CUR, curSol.data()));
        if (!stationary_ && !prevUploaded_) {
            ctx_->check(dmx_vec_upload(ctx_->get(), DMX_VEC_PREV, prevSol_->data()));
            prevUploaded_ = true;
        }
    }

    std::shared_ptr<Context> ctx_;
    bool stationary_;
    const SolutionVector* prevSol_ = nullptr;
    bool prevUploaded_ = false, hostJacobianValid_ = false, hostResidualValid_ = false;
    int numEq_ = 0;
    std::size_t nnzb_ = 0;
    dmx_options options_{};
    JacobianMatrix jac_;
    ResidualType res_;
};

// =====================================================================================================================
class GpuILUBiCGSTABSolver {
public:
    //! LinearSolverParameters defaults (linearsolverparameters.hh:56-73): maxit 250, reduction 1e-13 -- NewtonSolver
    //! overrides the reduction with LinearSolver.ResidualReduction = 1e-6 through setResidualReduction (newtonsolver.hh:232)
    explicit GpuILUBiCGSTABSolver(std::shared_ptr<Context> ctx) : GpuILUBiCGSTABSolver(std::move(ctx), DMX_SOLVER_BICGSTAB, 0) {}
    virtual ~GpuILUBiCGSTABSolver() = default;

    //! istlsolvers.hh:273: host matrix and vectors; x is the initial guess and the result
    IstlSolverResult solve(BCRSMatrix& A, BlockVector& x, BlockVector& b)
    {
        ensurePattern_(A);
        select();
        IstlSolverResult r;
        const int st = ctx_->check(dmx_linear_solve_host(ctx_->get(), A.values.data(), x.data(), b.data(), reduction_, maxIter_, precond_,
                                                          &r.iterations, &r.reduction),
                                   true);
        r.converged = (st == DMX_STATUS_OK);
        return r;
    }
    //! device-resident form: the system the assembler just assembled (no host copies of A and b); x receives deltaU
    IstlSolverResult solve(GpuFVAssembler& assembler, BlockVector& x)
    {
        dmx_ctx* c = ctx_->get();
        select();
        ctx_->check(dmx_vec_upload(c, DMX_VEC_DELTA, x.data()));
        IstlSolverResult r;
        const int st = ctx_->check(dmx_linear_solve(c, reduction_, maxIter_, precond_, &r.iterations, &r.reduction), true);
        r.converged = (st == DMX_STATUS_OK);
        ctx_->check(dmx_vec_download(c, DMX_VEC_DELTA, x.data()));
        (void)assembler;
        return r;
    }
    //! istlsolvers.hh:306-337: (owner-masked, all-reduced) 2-norm
    double norm(const BlockVector& v)
    {
        double out = 0.0;
        ctx_->check(dmx_vec_upload(ctx_->get(), DMX_VEC_WORK1, v.data()));
        ctx_->check(dmx_norm2(ctx_->get(), DMX_VEC_WORK1, &out));
        return out;
    }
    void setResidualReduction(double r) { reduction_ = r; }       // :350
    void setMaxIter(std::size_t i) { maxIter_ = static_cast<int>(i); }
    void setPreconditioner(int p) { precond_ = p; }               // DMX_PRECOND_ILU0 (default) or DMX_PRECOND_BLOCKJACOBI
    double residualReduction() const { return reduction_; }
    int maxIter() const { return maxIter_; }
    int preconditioner() const { return precond_; }
    //! makes this object's Krylov method the one the context runs (a property of the context, dmx_set_linear_solver); called
    //! before every solve so that several solver objects can share a context
    void select() const { ctx_->check(dmx_set_linear_solver(ctx_->get(), solver_, restart_)); }
    virtual std::string name() const { return "ILU0 preconditioned BiCGSTAB solver (B200)"; }

protected:
    GpuILUBiCGSTABSolver(std::shared_ptr<Context> ctx, int solver, int restart) : ctx_(std::move(ctx)), solver_(solver), restart_(restart)
    {
        select();
    }

private:
    void ensurePattern_(const BCRSMatrix& A)
    {
        dmx_ctx* c = ctx_->get();
        if (dmx_num_cells(c) == A.n && dmx_num_eq(c) == A.b && dmx_nnz_blocks(c) == static_cast<long long>(A.colidx.size())) return;
        ctx_->check(dmx_bcrs_pattern(c, A.n, A.b, A.rowptr.data(), A.colidx.data()));
    }
    std::shared_ptr<Context> ctx_;
    int solver_ = DMX_SOLVER_BICGSTAB, restart_ = 0;
    double reduction_ = 1e-13;
    int maxIter_ = 250, precond_ = DMX_PRECOND_ILU0;
};

// =====================================================================================================================
//! ILURestartedGMResIstlSolver (linear/istlsolvers.hh:660-667; what test/porousmediumflow/2p/incompressible/main.cc:134 uses):
//! same interface, Dune::RestartedGMResSolver instead of BiCGSTAB; restart = LinearSolver.GMResRestart (default 10)
class GpuILURestartedGMResSolver : public GpuILUBiCGSTABSolver {
public:
    explicit GpuILURestartedGMResSolver(std::shared_ptr<Context> ctx, int restart = 10)
    : GpuILUBiCGSTABSolver(std::move(ctx), DMX_SOLVER_RESTARTED_GMRES, restart) {}
    std::string name() const override { return "ILU0 preconditioned restarted GMRes solver (B200)"; }
};

//! SSORCGIstlSolver (linear/istlsolvers.hh:701-714; the linear solver of test/porousmediumflow/1p/incompressible/main.cc):
//! Dune::CGSolver preconditioned with one SeqSSOR iteration
class GpuSSORCGSolver : public GpuILUBiCGSTABSolver {
public:
    explicit GpuSSORCGSolver(std::shared_ptr<Context> ctx) : GpuILUBiCGSTABSolver(std::move(ctx), DMX_SOLVER_CG, 0) { setPreconditioner(DMX_PRECOND_SSOR); }
    std::string name() const override { return "SSOR preconditioned CG solver (B200)"; }
};
//! SSORBiCGSTABIstlSolver (linear/istlsolvers.hh:686-699)
class GpuSSORBiCGSTABSolver : public GpuILUBiCGSTABSolver {
public:
    explicit GpuSSORBiCGSTABSolver(std::shared_ptr<Context> ctx) : GpuILUBiCGSTABSolver(std::move(ctx), DMX_SOLVER_BICGSTAB, 0) { setPreconditioner(DMX_PRECOND_SSOR); }
    std::string name() const override { return "SSOR preconditioned BiCGSTAB solver (B200)"; }
};

//! AMGBiCGSTABIstlSolver (linear/istlsolvers.hh:716-736): BiCGSTAB preconditioned with one AMG V-cycle (dune-istl's default
//! cycle parameters, see dmx_amg_params); setAmgParams stands for the LinearSolver.Preconditioner.* keys Dune::AMGCreator reads
class GpuAMGBiCGSTABSolver : public GpuILUBiCGSTABSolver {
public:
    explicit GpuAMGBiCGSTABSolver(std::shared_ptr<Context> ctx) : GpuILUBiCGSTABSolver(ctx, DMX_SOLVER_BICGSTAB, 0), amgCtx_(std::move(ctx))
    {
        setPreconditioner(DMX_PRECOND_AMG);
    }
    void setAmgParams(const dmx_amg_params& p) { amgCtx_->check(dmx_set_amg_params(amgCtx_->get(), &p)); }
    std::string name() const override { return "AMG preconditioned BiCGSTAB solver (B200)"; }
private:
    std::shared_ptr<Context> amgCtx_;
};
//! AMGCGIstlSolver (linear/istlsolvers.hh:738-757)
class GpuAMGCGSolver : public GpuILUBiCGSTABSolver {
public:
    explicit GpuAMGCGSolver(std::shared_ptr<Context> ctx) : GpuILUBiCGSTABSolver(std::move(ctx), DMX_SOLVER_CG, 0) { setPreconditioner(DMX_PRECOND_AMG); }
    std::string name() const override { return "AMG preconditioned CG solver (B200)"; }
};

// =====================================================================================================================
class GpuNewtonSolver {
public:
    GpuNewtonSolver(std::shared_ptr<GpuFVAssembler> assembler, std::shared_ptr<GpuILUBiCGSTABSolver> linearSolver)
    : assembler_(std::move(assembler)), linearSolver_(std::move(linearSolver))
    {
        dmx_default_newton_params(&params_);      // newtonsolver.hh:1213-1247
        // the NewtonSolver constructor sets the linear solver's reduction to LinearSolver.ResidualReduction (default 1e-6), :232
        linearSolver_->setResidualReduction(params_.lin_reduction);
    }
    void setMaxRelativeShift(double s) { params_.max_relative_shift = s; }
    void setMinSteps(int n) { params_.min_steps = n; }
    void setMaxSteps(int n) { params_.max_steps = n; }
    void setTargetSteps(int n) { targetSteps_ = n; }
    void setMaxAbsoluteResidual(double r) { params_.max_absolute_residual = r; }                 // newtonsolver.hh:259
    void setResidualReduction(double r) { params_.residual_reduction = r; }                      // :268 (Newton.ResidualReduction)
    void setUseLineSearch(bool v = true) { params_.use_line_search = v ? 1 : 0; }               // :815
    //! Newton.EnableShiftCriterion / EnableResidualCriterion / EnableAbsoluteResidualCriterion /
    //! SatisfyResidualAndShiftCriterion (:1220-1223)
    void setConvergenceCriteria(bool shift, bool residual, bool absoluteResidual = false, bool satisfyBoth = false)
    {
        params_.enable_shift_criterion = shift ? 1 : 0;
        params_.enable_residual_criterion = residual ? 1 : 0;
        params_.enable_absolute_residual_criterion = absoluteResidual ? 1 : 0;
        params_.satisfy_residual_and_shift = satisfyBoth ? 1 : 0;
    }
    //! LinearSolver.ResidualReduction as set by the NewtonSolver constructor (:232) / LinearSolver.MaxIterations
    void setLinearResidualReduction(double r) { linearSolver_->setResidualReduction(r); }
    void setLinearMaxIterations(int n) { linearSolver_->setMaxIter(static_cast<std::size_t>(n)); }
    GpuILUBiCGSTABSolver& linearSolver() { return *linearSolver_; }
    GpuFVAssembler& assembler() { return *assembler_; }

    //! NewtonSolver::solve(vars) at fixed dt (newtonsolver.hh:362-372): throws NumericalProblem if not converged.
    //! The whole loop (assemble, solve, update, shift) runs on the device; u is uploaded once and downloaded once.
    void solve(BlockVector& u)
    {
        const auto& ctx = assembler_->context();
        const double* prev = assembler_->isStationaryProblem() ? nullptr : assembler_->prevSol().data();
        // the linear solve inside the device loop is the one `linearSolver_` describes: Krylov method, preconditioner,
        // reduction and iteration limit (solveLinearSystem calls linearSolver().solve, newtonsolver.hh:1201-1210)
        linearSolver_->select();
        params_.preconditioner = linearSolver_->preconditioner();
        params_.lin_reduction = linearSolver_->residualReduction();
        params_.lin_maxit = linearSolver_->maxIter();
        const int st = ctx->check(dmx_newton_solve_host(ctx->get(), u.data(), prev, &params_, &report_), true);
        if (st != DMX_STATUS_OK) throw NumericalProblem("Newton solver didn't converge after " + std::to_string(report_.newton_iterations) + " iterations");
    }
    //! NewtonSolver::solve(vars, timeLoop) (newtonsolver.hh:309-355): on failure reset to prevSol and halve dt, at most
    //! maxTimeStepDivisions times; returns the dt that succeeded
    double solve(BlockVector& u, double dt, int maxTimeStepDivisions = 10, double retryFactor = 0.5)
    {
        for (int i = 0; i <= maxTimeStepDivisions; ++i) {
            assembler_->setTimeStepSize(dt);
            try {
                solve(u);
                return dt;
            } catch (const NumericalProblem&) {
                if (i == maxTimeStepDivisions) break;
                u = assembler_->prevSol();
                assembler_->resetTimeStep(u);
                dt *= retryFactor;
            }
        }
        throw NumericalProblem("Newton solver didn't converge after " + std::to_string(maxTimeStepDivisions) + " time-step divisions");
    }
    //! newtonsolver.hh:784-798
    double suggestTimeStepSize(double oldTimeStep) const
    {
        const int n = report_.newton_iterations;
        if (n > targetSteps_) return oldTimeStep / (1.0 + static_cast<double>(n - targetSteps_) / targetSteps_);
        return oldTimeStep * (1.0 + static_cast<double>(targetSteps_ - n) / targetSteps_ / 1.2);
    }
    const dmx_newton_report& report() const { return report_; }

private:
    std::shared_ptr<GpuFVAssembler> assembler_;
    std::shared_ptr<GpuILUBiCGSTABSolver> linearSolver_;
    dmx_newton_params params_{};
    dmx_newton_report report_{};
    int targetSteps_ = 10;
};

} // namespace dumux_b200
#endif
