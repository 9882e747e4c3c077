// oracle/ref_solver.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md, ref_solver.h).
#include "ref_solver.h"
#include <limits>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace oracle {

void Problem::AddParameterBlock(double *data, int size, bool is_pose, int group) {
    if (index.count(data)) return;
    PB b; b.data = data; b.size = size; b.local = is_pose ? size - 1 : size; b.is_pose = is_pose;
    b.constant = false; b.group = group; b.col = -1; b.xoff = -1;
    index[data] = (int)pbs.size();
    pbs.push_back(b);
}
void Problem::SetParameterBlockConstant(double *data) { pbs[index.at(data)].constant = true; }
void Problem::AddResidualBlock(std::shared_ptr<CostFunction> cost, bool huber, const std::vector<double *> &params) {
    RB r; r.cost = cost; r.huber = huber;
    for (double *p : params) r.p.push_back(index.at(p));
    rbs.push_back(r);
}

namespace {

struct Program {
    Problem *P;
    std::vector<int> active;        // pb indices, e-blocks (group 0) first
    int num_e = 0, ne = 0, nf = 0, ncols = 0, nx = 0, nres = 0;
    std::vector<int> roff;          // residual offset per rb
    std::vector<std::vector<long>> joff;   // per rb, per slot: offset into jac (or -1)
    long jac_size = 0;
    std::vector<int> e_of_rb;       // e-block ordinal (0..num_e-1) or -1
    std::vector<std::vector<int>> rb_of_e;
    std::vector<int> rb_no_e;
    double huber_delta = 1.0;
};

void build_program(Problem &P, Program &G) {
    G.P = &P;
    for (int g = 0; g < 2; g++)
        for (size_t i = 0; i < P.pbs.size(); i++)
            if (!P.pbs[i].constant && P.pbs[i].group == g) G.active.push_back((int)i);
    int col = 0, xoff = 0;
    std::vector<int> e_ord(P.pbs.size(), -1);
    for (int idx : G.active) {
        Problem::PB &b = P.pbs[idx];
        if (b.group == 0) { e_ord[idx] = G.num_e++; G.ne += b.local; }
        b.col = col; col += b.local; b.xoff = xoff; xoff += b.size;
    }
    G.ncols = col; G.nf = col - G.ne; G.nx = xoff;
    G.rb_of_e.assign(G.num_e, {});
    int roff = 0; long joff = 0;
    for (size_t i = 0; i < P.rbs.size(); i++) {
        Problem::RB &rb = P.rbs[i];
        int nr = rb.cost->num_residuals;
        G.roff.push_back(roff); roff += nr;
        std::vector<long> jo; int e = -1;
        for (int pi : rb.p) {
            const Problem::PB &b = P.pbs[pi];
            if (b.constant) { jo.push_back(-1); continue; }
            jo.push_back(joff); joff += (long)nr * b.local;
            if (b.group == 0) e = e_ord[pi];
        }
        G.joff.push_back(jo); G.e_of_rb.push_back(e);
        if (e >= 0) G.rb_of_e[e].push_back((int)i); else G.rb_no_e.push_back((int)i);
    }
    G.nres = roff; G.jac_size = joff;
}

// Evaluator::Evaluate + ResidualBlock::Evaluate (local parameterization, then loss corrector).
bool evaluate(Program &G, const std::vector<double> &x, double &cost, std::vector<double> *res, std::vector<double> *jac) {
    Problem &P = *G.P;
    cost = 0;
    static thread_local std::vector<std::vector<double>> gbuf;
    static thread_local std::vector<double> rtmp;
    for (size_t i = 0; i < P.rbs.size(); i++) {
        Problem::RB &rb = P.rbs[i];
        int nr = rb.cost->num_residuals, np = (int)rb.p.size();
        const double *params[32]; double *jptr[32];
        if ((int)gbuf.size() < np) gbuf.resize(np);
        for (int k = 0; k < np; k++) {
            const Problem::PB &b = P.pbs[rb.p[k]];
            params[k] = b.constant ? b.data : &x[b.xoff];
            if (jac && !b.constant) {
                if ((int)gbuf[k].size() < nr * b.size) gbuf[k].resize((size_t)nr * b.size);
                jptr[k] = gbuf[k].data();
            } else jptr[k] = nullptr;
        }
        if ((int)rtmp.size() < nr) rtmp.resize(nr);
        double *r = res ? &(*res)[G.roff[i]] : rtmp.data();
        if (!rb.cost->Evaluate(params, r, jac ? jptr : nullptr)) return false;
        double sq = 0; for (int k = 0; k < nr; k++) sq += r[k] * r[k];
        double sqrt_rho1 = 1.0, residual_scaling = 1.0, alpha_sq_norm = 0.0;
        if (!rb.huber) cost += 0.5 * sq;
        else {
            double rho[3]; huber_loss(G.huber_delta, sq, rho);
            cost += 0.5 * rho[0];
            sqrt_rho1 = std::sqrt(rho[1]);                       // Corrector::Corrector (corrector.cc)
            if (sq == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
            else {
                const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
                const double alpha = 1.0 - std::sqrt(D);
                residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / sq;
            }
        }
        if (jac) {
            for (int k = 0; k < np; k++) {
                const Problem::PB &b = P.pbs[rb.p[k]];
                if (b.constant) continue;
                double *out = &(*jac)[G.joff[i][k]];
                const double *gj = gbuf[k].data();
                // global -> local: J_global * ComputeJacobian ([I6;0] for poses, identity otherwise)
                for (int rr = 0; rr < nr; rr++) for (int c = 0; c < b.local; c++) out[rr * b.local + c] = gj[rr * b.size + c];
                if (rb.huber) {                                    // Corrector::CorrectJacobian
                    if (alpha_sq_norm == 0.0) { for (int t = 0; t < nr * b.local; t++) out[t] *= sqrt_rho1; }
                    else {
                        for (int c = 0; c < b.local; c++) {
                            double rtj = 0; for (int rr = 0; rr < nr; rr++) rtj += out[rr * b.local + c] * r[rr];
                            for (int rr = 0; rr < nr; rr++) out[rr * b.local + c] = sqrt_rho1 * (out[rr * b.local + c] - alpha_sq_norm * r[rr] * rtj);
                        }
                    }
                }
            }
        }
        if (rb.huber && res) for (int k = 0; k < nr; k++) r[k] *= residual_scaling;   // CorrectResiduals
    }
    return std::isfinite(cost);
}

// y += J * v  (J as stored, i.e. scaled once ScaleColumns has run)
void jac_right_multiply(const Program &G, const std::vector<double> &jac, const double *v, double *y) {
    const Problem &P = *G.P;
    for (size_t i = 0; i < P.rbs.size(); i++) {
        const Problem::RB &rb = P.rbs[i]; int nr = rb.cost->num_residuals;
        for (size_t k = 0; k < rb.p.size(); k++) {
            long jo = G.joff[i][k]; if (jo < 0) continue;
            const Problem::PB &b = P.pbs[rb.p[k]];
            for (int rr = 0; rr < nr; rr++) { double s = 0; for (int c = 0; c < b.local; c++) s += jac[jo + rr * b.local + c] * v[b.col + c]; y[G.roff[i] + rr] += s; }
        }
    }
}
// y += J^T * r
void jac_left_multiply(const Program &G, const std::vector<double> &jac, const double *r, double *y) {
    const Problem &P = *G.P;
    for (size_t i = 0; i < P.rbs.size(); i++) {
        const Problem::RB &rb = P.rbs[i]; int nr = rb.cost->num_residuals;
        for (size_t k = 0; k < rb.p.size(); k++) {
            long jo = G.joff[i][k]; if (jo < 0) continue;
            const Problem::PB &b = P.pbs[rb.p[k]];
            for (int rr = 0; rr < nr; rr++) { double rv = r[G.roff[i] + rr]; for (int c = 0; c < b.local; c++) y[b.col + c] += jac[jo + rr * b.local + c] * rv; }
        }
    }
}
void squared_column_norm(const Program &G, const std::vector<double> &jac, double *out) {
    const Problem &P = *G.P;
    for (int c = 0; c < G.ncols; c++) out[c] = 0;
    for (size_t i = 0; i < P.rbs.size(); i++) {
        const Problem::RB &rb = P.rbs[i]; int nr = rb.cost->num_residuals;
        for (size_t k = 0; k < rb.p.size(); k++) {
            long jo = G.joff[i][k]; if (jo < 0) continue;
            const Problem::PB &b = P.pbs[rb.p[k]];
            for (int rr = 0; rr < nr; rr++) for (int c = 0; c < b.local; c++) { double v = jac[jo + rr * b.local + c]; out[b.col + c] += v * v; }
        }
    }
}
void scale_columns(const Program &G, std::vector<double> &jac, const double *s) {
    const Problem &P = *G.P;
    for (size_t i = 0; i < P.rbs.size(); i++) {
        const Problem::RB &rb = P.rbs[i]; int nr = rb.cost->num_residuals;
        for (size_t k = 0; k < rb.p.size(); k++) {
            long jo = G.joff[i][k]; if (jo < 0) continue;
            const Problem::PB &b = P.pbs[rb.p[k]];
            for (int rr = 0; rr < nr; rr++) for (int c = 0; c < b.local; c++) jac[jo + rr * b.local + c] *= s[b.col + c];
        }
    }
}

// SchurComplementSolver (DENSE_SCHUR): solve (J^T J + D^T D) y = J^T b, e-blocks = group 0,
// reduced system factored by dense Cholesky (LLT).  Returns false == LINEAR_SOLVER_FAILURE.
bool schur_solve(const Program &G, const std::vector<double> &jac, const std::vector<double> &b, const double *D, double *y) {
    const Problem &P = *G.P;
    const int nf = G.nf, ne = G.ne;
    Mat lhs(nf, nf);
    std::vector<double> rhs(nf, 0.0);
    for (int c = 0; c < nf; c++) lhs(c, c) = D[ne + c] * D[ne + c];
    auto add_ftf = [&](int i) {        // F^T F and F^T b of one residual block
        const Problem::RB &rb = P.rbs[i]; int nr = rb.cost->num_residuals;
        for (size_t ka = 0; ka < rb.p.size(); ka++) {
            long ja = G.joff[i][ka]; if (ja < 0) continue;
            const Problem::PB &ba = P.pbs[rb.p[ka]]; if (ba.group == 0) continue;
            int ca = ba.col - ne;
            for (int rr = 0; rr < nr; rr++) { double bv = b[G.roff[i] + rr]; for (int c = 0; c < ba.local; c++) rhs[ca + c] += jac[ja + rr * ba.local + c] * bv; }
            for (size_t kb = ka; kb < rb.p.size(); kb++) {
                long jb = G.joff[i][kb]; if (jb < 0) continue;
                const Problem::PB &bb = P.pbs[rb.p[kb]]; if (bb.group == 0) continue;
                int cb = bb.col - ne;
                for (int c1 = 0; c1 < ba.local; c1++) for (int c2 = 0; c2 < bb.local; c2++) {
                    double s = 0; for (int rr = 0; rr < nr; rr++) s += jac[ja + rr * ba.local + c1] * jac[jb + rr * bb.local + c2];
                    lhs(ca + c1, cb + c2) += s;
                    if (ka != kb) lhs(cb + c2, ca + c1) += s;
                }
            }
        }
    };
    for (int i : G.rb_no_e) add_ftf(i);
    // chunks: one per e-block (size-1 e-blocks: the inverse depths)
    std::vector<double> W((size_t)G.num_e * nf, 0.0), ete(G.num_e, 0.0), ge(G.num_e, 0.0);
    std::vector<int> ecol(G.num_e, 0);
    for (int e = 0; e < G.num_e; e++) {
        double *w = &W[(size_t)e * nf];
        int lo = nf, hi = 0;
        for (int i : G.rb_of_e[e]) {
            add_ftf(i);
            const Problem::RB &rb = P.rbs[i]; int nr = rb.cost->num_residuals;
            long je = -1; int cole = 0;
            for (size_t k = 0; k < rb.p.size(); k++) if (G.joff[i][k] >= 0 && P.pbs[rb.p[k]].group == 0) { je = G.joff[i][k]; cole = P.pbs[rb.p[k]].col; }
            ecol[e] = cole;
            for (int rr = 0; rr < nr; rr++) { ete[e] += jac[je + rr] * jac[je + rr]; ge[e] += jac[je + rr] * b[G.roff[i] + rr]; }
            for (size_t k = 0; k < rb.p.size(); k++) {
                long jo = G.joff[i][k]; if (jo < 0) continue;
                const Problem::PB &bb = P.pbs[rb.p[k]]; if (bb.group == 0) continue;
                int cb = bb.col - ne;
                for (int c = 0; c < bb.local; c++) { double s = 0; for (int rr = 0; rr < nr; rr++) s += jac[jo + rr * bb.local + c] * jac[je + rr]; w[cb + c] += s; }
                lo = std::min(lo, cb); hi = std::max(hi, cb + bb.local);
            }
        }
        ete[e] += D[ecol[e]] * D[ecol[e]];
        double inv = 1.0 / ete[e];
        for (int a = lo; a < hi; a++) {
            double wa = w[a] * inv; if (wa == 0.0) continue;
            rhs[a] -= wa * ge[e];
            for (int c = lo; c < hi; c++) lhs(a, c) -= wa * w[c];
        }
    }
    if (!cholesky_lower(lhs)) return false;
    chol_solve_inplace(lhs, rhs.data());
    for (int c = 0; c < nf; c++) y[ne + c] = rhs[c];
    for (int e = 0; e < G.num_e; e++) {       // back substitution
        const double *w = &W[(size_t)e * nf];
        double s = ge[e]; for (int c = 0; c < nf; c++) s -= w[c] * rhs[c];
        y[ecol[e]] = s / ete[e];
    }
    for (int c = 0; c < G.ncols; c++) if (!std::isfinite(y[c])) return false;
    return true;
}

// DoglegStrategy (dogleg_strategy.cc), TRADITIONAL_DOGLEG
struct Dogleg {
    double radius, max_radius, min_diagonal = 1e-6, max_diagonal = 1e32;
    double mu = 1e-8, min_mu = 1e-8, max_mu = 1.0, mu_increase_factor = 10.0;
    double increase_threshold = 0.75, decrease_threshold = 0.25;
    double dogleg_step_norm = 0, alpha = 0;
    bool reuse = false;
    int gn_attempts = 0, test_fail_factorizations = 0;      // CERB_TEST_FAIL_FACTORIZATIONS (environment): report the first k linear solves as LINEAR_SOLVER_FAILURE
    std::vector<double> diagonal, gradient, gauss_newton_step;

    // returns 0 SUCCESS, 1 FAILURE
    int ComputeStep(const Program &G, const std::vector<double> &jac, const std::vector<double> &residuals, double *step) {
        int n = G.ncols;
        if (reuse) { ComputeTraditionalDoglegStep(n, step); return 0; }
        reuse = true;
        diagonal.assign(n, 0.0); gradient.assign(n, 0.0); gauss_newton_step.assign(n, 0.0);
        squared_column_norm(G, jac, diagonal.data());
        for (int i = 0; i < n; i++) diagonal[i] = std::sqrt(std::min(std::max(diagonal[i], min_diagonal), max_diagonal));
        // ComputeGradient
        jac_left_multiply(G, jac, residuals.data(), gradient.data());
        for (int i = 0; i < n; i++) gradient[i] /= diagonal[i];
        // ComputeCauchyPoint
        std::vector<double> Jg(G.nres, 0.0), sg(n);
        for (int i = 0; i < n; i++) sg[i] = gradient[i] / diagonal[i];
        jac_right_multiply(G, jac, sg.data(), Jg.data());
        double g2 = 0, jg2 = 0; for (double v : gradient) g2 += v * v; for (double v : Jg) jg2 += v * v;
        alpha = g2 / jg2;
        // ComputeGaussNewtonStep
        bool ok = false;
        while (mu < max_mu) {
            std::vector<double> lm(n); for (int i = 0; i < n; i++) lm[i] = diagonal[i] * std::sqrt(mu);
            bool solved = schur_solve(G, jac, residuals, lm.data(), gauss_newton_step.data());
            if (gn_attempts++ < test_fail_factorizations) solved = false;         // fault injection (tests/test_solver_failure.py)
            if (solved) { ok = true; break; }
            mu *= mu_increase_factor;
        }
        if (!ok) return 1;
        for (int i = 0; i < n; i++) gauss_newton_step[i] *= -diagonal[i];
        ComputeTraditionalDoglegStep(n, step);
        return 0;
    }
    void ComputeTraditionalDoglegStep(int n, double *dogleg) {
        double gradient_norm = 0, gauss_newton_norm = 0;
        for (int i = 0; i < n; i++) { gradient_norm += gradient[i] * gradient[i]; gauss_newton_norm += gauss_newton_step[i] * gauss_newton_step[i]; }
        gradient_norm = std::sqrt(gradient_norm); gauss_newton_norm = std::sqrt(gauss_newton_norm);
        if (gauss_newton_norm <= radius) {                                   // case 1
            for (int i = 0; i < n; i++) dogleg[i] = gauss_newton_step[i] / diagonal[i];
            dogleg_step_norm = gauss_newton_norm; return;
        }
        if (gradient_norm * alpha >= radius) {                               // case 2
            for (int i = 0; i < n; i++) dogleg[i] = -(radius / gradient_norm) * gradient[i] / diagonal[i];
            dogleg_step_norm = radius; return;
        }
        double gdotgn = 0; for (int i = 0; i < n; i++) gdotgn += gradient[i] * gauss_newton_step[i];   // case 3
        const double b_dot_a = -alpha * gdotgn;
        const double a_squared_norm = std::pow(alpha * gradient_norm, 2.0);
        const double b_minus_a_squared_norm = a_squared_norm - 2 * b_dot_a + std::pow(gauss_newton_norm, 2);
        const double c = b_dot_a - a_squared_norm;
        const double d = std::sqrt(c * c + b_minus_a_squared_norm * (std::pow(radius, 2.0) - a_squared_norm));
        double beta = (c <= 0) ? (d - c) / b_minus_a_squared_norm : (radius * radius - a_squared_norm) / (d + c);
        double nrm = 0;
        for (int i = 0; i < n; i++) { double v = (-alpha * (1.0 - beta)) * gradient[i] + beta * gauss_newton_step[i]; nrm += v * v; dogleg[i] = v / diagonal[i]; }
        dogleg_step_norm = std::sqrt(nrm);
    }
    void StepAccepted(double q) {
        if (q < decrease_threshold) radius *= 0.5;
        if (q > increase_threshold) radius = std::max(radius, 3.0 * dogleg_step_norm);
        mu = std::max(min_mu, 2.0 * mu / mu_increase_factor);
        reuse = false;
    }
    void StepRejected(double) { radius *= 0.5; reuse = true; }
    void StepIsInvalid() { mu *= mu_increase_factor; reuse = false; }
};

void plus(const Program &G, const std::vector<double> &x, const double *delta, std::vector<double> &out) {
    const Problem &P = *G.P;
    out.resize(x.size());
    for (int idx : G.active) {
        const Problem::PB &b = P.pbs[idx];
        if (b.is_pose) pose_plus(&x[b.xoff], delta + b.col, &out[b.xoff]);
        else for (int k = 0; k < b.size; k++) out[b.xoff + k] = x[b.xoff + k] + delta[b.col + k];
    }
}
double vnorm(const std::vector<double> &v) { double s = 0; for (double a : v) s += a * a; return std::sqrt(s); }

}  // namespace

// TrustRegionMinimizer::Minimize (trust_region_minimizer.cc, Ceres 1.14.0)
void Solve(const SolverOptions &opt, Problem &problem, SolverSummary &summary) {
    Program G; build_program(problem, G); G.huber_delta = opt.huber_delta;
    const int n = G.ncols;
    std::vector<double> x(G.nx), candidate_x, residuals(G.nres), jac(G.jac_size), gradient(n), jacobian_scaling(n, 1.0);
    std::vector<double> trust_region_step(n), delta(n);
    for (int idx : G.active) { const Problem::PB &b = problem.pbs[idx]; for (int k = 0; k < b.size; k++) x[b.xoff + k] = b.data[k]; }
    double x_norm = vnorm(x), x_cost = 0, candidate_cost = 0, model_cost_change = 0;
    Dogleg strategy; strategy.radius = opt.initial_trust_region_radius; strategy.max_radius = opt.max_trust_region_radius;
    if (const char *e = getenv("CERB_TEST_FAIL_FACTORIZATIONS")) strategy.test_fail_factorizations = atoi(e);     // same hooks as the product library
    if (const char *e = getenv("CERB_TEST_INITIAL_MU")) strategy.mu = atof(e);
    int iteration = 0, num_consecutive_invalid_steps = 0;
    bool step_is_successful = false;
    double gradient_max_norm = 0;
    summary = SolverSummary();

    auto EvaluateGradientAndJacobian = [&]() -> bool {
        if (!evaluate(G, x, x_cost, &residuals, &jac)) return false;
        std::fill(gradient.begin(), gradient.end(), 0.0);
        jac_left_multiply(G, jac, residuals.data(), gradient.data());      // unscaled gradient
        if (iteration == 0) {
            squared_column_norm(G, jac, jacobian_scaling.data());
            summary.gradient0 = gradient; summary.jtj_diag0 = jacobian_scaling;
            for (int i = 0; i < n; i++) jacobian_scaling[i] = 1.0 / (1.0 + std::sqrt(jacobian_scaling[i]));
        }
        scale_columns(G, jac, jacobian_scaling.data());
        gradient_max_norm = 0; for (double g : gradient) gradient_max_norm = std::max(gradient_max_norm, std::fabs(g));
        return true;
    };
    auto write_back = [&]() {
        for (int idx : G.active) { const Problem::PB &b = problem.pbs[idx]; for (int k = 0; k < b.size; k++) b.data[k] = x[b.xoff + k]; }
        summary.final_cost = x_cost;
    };

    // IterationZero
    if (!EvaluateGradientAndJacobian()) { summary.termination = 2; summary.initial_cost = summary.final_cost = x_cost; return; }
    summary.initial_cost = x_cost;

    while (true) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        if (step_is_successful) summary.num_successful_steps++;
        summary.iterations = iteration;
        if (iteration >= opt.max_num_iterations) { summary.termination = 1; break; }
        if (gradient_max_norm <= opt.gradient_tolerance) { summary.termination = 0; break; }
        if (strategy.radius <= opt.min_trust_region_radius) { summary.termination = 0; break; }
        iteration++;
        step_is_successful = false;

        // ComputeTrustRegionStep
        bool step_is_valid = false;
        int st = strategy.ComputeStep(G, jac, residuals, trust_region_step.data());
        if (st == 0) {
            std::vector<double> model_residuals(G.nres, 0.0);
            jac_right_multiply(G, jac, trust_region_step.data(), model_residuals.data());
            double s = 0; for (int i = 0; i < G.nres; i++) s += model_residuals[i] * (residuals[i] + model_residuals[i] / 2.0);
            model_cost_change = -s;
            step_is_valid = model_cost_change > 0.0;
            if (step_is_valid) { for (int i = 0; i < n; i++) delta[i] = trust_region_step[i] * jacobian_scaling[i]; num_consecutive_invalid_steps = 0; }
        }
        if (!step_is_valid) {   // HandleInvalidStep
            if (++num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) { summary.termination = 2; summary.iterations = iteration; break; }
            strategy.StepIsInvalid();
            continue;
        }
        // ComputeCandidatePointAndEvaluateCost
        plus(G, x, delta.data(), candidate_x);
        if (!evaluate(G, candidate_x, candidate_cost, nullptr, nullptr)) candidate_cost = std::numeric_limits<double>::max();
        // ParameterToleranceReached
        double step_norm = 0; for (int i = 0; i < G.nx; i++) step_norm += (x[i] - candidate_x[i]) * (x[i] - candidate_x[i]);
        step_norm = std::sqrt(step_norm);
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { summary.termination = 0; summary.iterations = iteration; break; }
        // FunctionToleranceReached
        double cost_change = x_cost - candidate_cost;
        if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { summary.termination = 0; summary.iterations = iteration; break; }
        // IsStepSuccessful (monotonic TrustRegionStepEvaluator)
        double relative_decrease = cost_change / model_cost_change;
        if (getenv("ORACLE_TRACE")) fprintf(stderr, "it %2d cost %.9e cand %.9e model_change %.3e rel %.4f radius %.3e step_norm %.3e mu %.1e gmax %.3e\n",
            iteration, x_cost, candidate_cost, model_cost_change, relative_decrease, strategy.radius, step_norm, strategy.mu, gradient_max_norm);
        if (relative_decrease > opt.min_relative_decrease) {   // HandleSuccessfulStep
            x = candidate_x; x_norm = vnorm(x);
            if (!EvaluateGradientAndJacobian()) { summary.termination = 2; summary.iterations = iteration; break; }
            step_is_successful = true;
            strategy.StepAccepted(relative_decrease);
        } else {
            strategy.StepRejected(relative_decrease);
        }
    }
    write_back();
}

}  // namespace oracle
