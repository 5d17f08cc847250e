// The duck-typed contract DuMux's NewtonSolver imposes on its assembler and linear solver, as documented by the mocks of
// test/nonlinear/newton/test_newton.cc:31-78 -- and nothing more:
//   Assembler:     types Scalar, ResidualType, JacobianMatrix, SolutionVector, Variables; setLinearSystem();
//                  assembleResidual(sol); assembleJacobianAndResidual(sol); jacobian(); residual()
//   LinearSolver:  setResidualReduction(double); solve(A, x, b) convertible to bool; norm(residual)
// MiniNewton below is a Newton loop written against exactly these members (the call sequence of NewtonSolver::solveImpl_,
// nonlinear/newtonsolver.hh:976-1072: assembleLinearSystem :469, solveLinearSystem :488-524 incl. norm(residual) :495,
// newtonUpdate :543-557).  It is instantiated (1) with the reference's own mock classes, restated, and run: x^2 - 5 = 0 to
// 1e-13 as in the reference test; (2) with GpuFVAssembler / the Gpu*Solver classes of include/dumux_b200.hpp -- if that
// compiles, the drop-in classes satisfy the contract NewtonSolver needs.  (2) only runs when a GPU is visible (argv[1] = "gpu").
#include <cmath>
#include <cstdio>
#include <memory>
#include <string>
#include <type_traits>

#include "dumux_b200.hpp"

namespace contract {

template <class T> double maxShift(const T& a, const T& b);
template <> double maxShift<double>(const double& a, const double& b) { return std::fabs(a - b) / std::fmax(1.0, std::fabs(a + b) * 0.5); }
template <> double maxShift<dumux_b200::BlockVector>(const dumux_b200::BlockVector& a, const dumux_b200::BlockVector& b)
{
    double s = 0.0;
    const double *pa = a.data(), *pb = b.data();
    for (std::size_t i = 0; i < a.size() * a.blockSize(); ++i) s = std::fmax(s, std::fabs(pa[i] - pb[i]) / std::fmax(1.0, std::fabs(pa[i] + pb[i]) * 0.5));
    return s;
}

template <class Assembler, class LinearSolver>
class MiniNewton {
public:
    using Scalar = typename Assembler::Scalar;
    using SolutionVector = typename Assembler::SolutionVector;
    using ResidualType = typename Assembler::ResidualType;
    using JacobianMatrix = typename Assembler::JacobianMatrix;
    using Variables = typename Assembler::Variables;
    static_assert(std::is_same<Variables, SolutionVector>::value, "assemblers that do not export grid variables: Variables = SolutionVector");

    MiniNewton(std::shared_ptr<Assembler> a, std::shared_ptr<LinearSolver> ls) : assembler_(std::move(a)), linearSolver_(std::move(ls))
    {
        linearSolver_->setResidualReduction(1e-6);                 // newtonsolver.hh:232
        assembler_->setLinearSystem();
    }
    int solve(SolutionVector& u, int maxSteps = 18, double maxRelativeShift = 1e-8)
    {
        int steps = 0;
        double shift = 1.0;
        while (steps < 2 || (shift > maxRelativeShift && steps < maxSteps)) {
            const SolutionVector uLast = u;
            assembler_->assembleJacobianAndResidual(u);
            const double initialNorm = linearSolver_->norm(assembler_->residual());
            (void)initialNorm;
            SolutionVector deltaU = u;
            deltaU = 0.0;
            JacobianMatrix& A = assembler_->jacobian();
            ResidualType& b = assembler_->residual();
            const bool converged = static_cast<bool>(linearSolver_->solve(A, deltaU, b));
            if (!converged) return -1;
            u -= deltaU;
            shift = maxShift(u, uLast);
            assembler_->assembleResidual(u);
            ++steps;
        }
        return steps;
    }

private:
    std::shared_ptr<Assembler> assembler_;
    std::shared_ptr<LinearSolver> linearSolver_;
};

// the reference's mocks, restated (test_newton.cc:31-78)
class MockScalarAssembler {
public:
    using Scalar = double;
    using ResidualType = Scalar;
    using JacobianMatrix = Scalar;
    using SolutionVector = Scalar;
    using Variables = Scalar;
    void setLinearSystem() {}
    void assembleResidual(const ResidualType& sol) { res_ = sol * sol - 5.0; }
    void assembleJacobianAndResidual(const ResidualType& sol) { assembleResidual(sol); jac_ = 2.0 * sol; }
    JacobianMatrix& jacobian() { return jac_; }
    ResidualType& residual() { return res_; }
private:
    JacobianMatrix jac_ = 0.0;
    ResidualType res_ = 0.0;
};
class MockScalarLinearSolver {
public:
    void setResidualReduction(double) {}
    bool solve(const double& A, double& x, const double& b) const { x = b / A; return true; }
    double norm(const double& residual) const { return std::fabs(residual); }
};

} // namespace contract

int main(int argc, char** argv)
{
    using namespace contract;
    {
        MiniNewton<MockScalarAssembler, MockScalarLinearSolver> newton(std::make_shared<MockScalarAssembler>(), std::make_shared<MockScalarLinearSolver>());
        double x = 0.1;
        const int steps = newton.solve(x, 50);
        if (steps < 0 || std::fabs(x - std::sqrt(5.0)) > 1e-13 * std::sqrt(5.0)) { std::fprintf(stderr, "mock Newton failed: %.17g\n", x); return 1; }
        std::printf("mock %d %.17g\n", steps, x);
    }
    if (argc > 1 && std::string(argv[1]) == "gpu") {
        // the same loop over the drop-in classes: 1p incompressible 6x6, every Gpu*Solver flavour instantiates the template
        using namespace dumux_b200;
        try {
            auto ctx = std::make_shared<Context>(0);
            ProblemData p;
            p.model = DMX_MODEL_1P; p.dim = 2; p.cells = {{6, 6, 1}};
            p.boundary[2].type.assign(6, DMX_BC_DIRICHLET); p.boundary[2].values.assign(6, 2.0e5);
            p.boundary[3].type.assign(6, DMX_BC_DIRICHLET); p.boundary[3].values.assign(6, 1.0e5);
            p.options.base_eps = 0.1; p.options.privar_magnitude[0] = 1e5;
            auto assembler = std::make_shared<GpuFVAssembler>(ctx, p);
            BlockVector x(assembler->numDofs(), 1, 0.0);
            MiniNewton<GpuFVAssembler, GpuILUBiCGSTABSolver> n1(assembler, std::make_shared<GpuILUBiCGSTABSolver>(ctx));
            const int s1 = n1.solve(x);
            BlockVector y(assembler->numDofs(), 1, 0.0);
            MiniNewton<GpuFVAssembler, GpuSSORCGSolver> n2(assembler, std::make_shared<GpuSSORCGSolver>(ctx));
            const int s2 = n2.solve(y);
            BlockVector z(assembler->numDofs(), 1, 0.0);
            MiniNewton<GpuFVAssembler, GpuILURestartedGMResSolver> n3(assembler, std::make_shared<GpuILURestartedGMResSolver>(ctx));
            const int s3 = n3.solve(z);
            std::printf("gpu %d %d %d %.17g %.17g %.17g\n", s1, s2, s3, x[0][0], y[0][0], z[0][0]);
        } catch (const std::exception& e) {
            std::fprintf(stderr, "error: %s\n", e.what());
            return 2;
        }
    }
    return 0;
}
