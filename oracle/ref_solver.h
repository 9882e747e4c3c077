// oracle/ref_solver.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Restatement of the part of ceres-solver 1.14.0 that Estimator::optimization() exercises
// (estimator.cpp:1059-1113 Problem set-up, :1221-1236 Options + Solve):
//   ceres::Problem {AddParameterBlock, SetParameterBlockConstant, AddResidualBlock},
//   TRUST_REGION minimizer, DOGLEG (TRADITIONAL_DOGLEG) strategy, DENSE_SCHUR linear solver with
//   dense Cholesky (Eigen LLT), Jacobi scaling, HuberLoss corrector, all other options default.
// ceres-solver is a THIRD-PARTY dependency that is NOT under /root/reference (pinned to tag 1.14.0 by
// .devcontainer/Dockerfile:69); its algorithm is restated from its published source
// (internal/ceres/trust_region_minimizer.cc, dogleg_strategy.cc, schur_complement_solver.cc,
// corrector.cc, loss_function.cc).  PARITY UNPINNED: no golden vectors exist for this boundary.
// The wall-clock cap (max_solver_time_in_seconds) is deliberately not restated: it makes the
// iterate returned machine dependent.
#pragma once
#include "ref_factors.h"
#include <map>

namespace oracle {

struct SolverOptions {
    int max_num_iterations = 12;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    int max_num_consecutive_invalid_steps = 5;
    double huber_delta = 1.0;
};

struct SolverSummary {
    int iterations = 0;             // trust-region iterations performed
    int num_successful_steps = 0;
    int termination = 1;            // 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE
    double initial_cost = 0, final_cost = 0;
    // probes for parity tests: linearisation at the initial point (tangent space, column order of
    // Problem::tangent_order()) -- gradient J^T r and diag(J^T J), unscaled, with the loss corrector applied
    std::vector<double> gradient0, jtj_diag0;
};

class Problem {
public:
    // group 0 = eliminated by the Schur complement (inverse depths), group 1 = the rest
    void AddParameterBlock(double *data, int size, bool is_pose, int group);
    void SetParameterBlockConstant(double *data);
    void AddResidualBlock(std::shared_ptr<CostFunction> cost, bool huber, const std::vector<double *> &params);

    struct PB { double *data; int size; int local; bool is_pose; bool constant; int group; int col; int xoff; };
    struct RB { std::shared_ptr<CostFunction> cost; bool huber; std::vector<int> p; };
    std::vector<PB> pbs;
    std::vector<RB> rbs;
    std::map<double *, int> index;
};

void Solve(const SolverOptions &opt, Problem &problem, SolverSummary &summary);

}  // namespace oracle
