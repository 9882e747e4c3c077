// oracle/ref_window.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Restatement of Estimator::optimization() (src/estimator/estimator.cpp:1054-1458) on top of the
// plain-C window structs of include/cerberus_b200.h, plus the extern "C" surface the python tests
// and bench.py's cpu_baseline / --impl reference leg bind with ctypes.  PARITY UNPINNED.
#include "../include/cerberus_b200.h"
#include "ref_solver.h"
#include "ref_preint.h"
#include <thread>
#include <atomic>
#include <cstdio>

using namespace oracle;

namespace {

LegPreintState to_state(const CerbIMULegPreint &p) {
    LegPreintState s;
    s.sum_dt = p.sum_dt;
    s.delta_p = V3(p.delta_p); s.delta_v = V3(p.delta_v);
    s.delta_q = Quat(p.delta_q[3], p.delta_q[0], p.delta_q[1], p.delta_q[2]);
    for (int j = 0; j < 4; j++) { s.delta_epsilon[j] = V3(p.delta_epsilon + 3 * j); s.linearized_rho[j] = p.linearized_rho[j]; }
    s.linearized_ba = V3(p.linearized_ba); s.linearized_bg = V3(p.linearized_bg);
    for (int r = 0; r < 31; r++) for (int c = 0; c < 31; c++) { s.jacobian(r, c) = p.jacobian[c * 31 + r]; s.covariance(r, c) = p.covariance[c * 31 + r]; }
    return s;
}
void from_state(const LegPreintState &s, CerbIMULegPreint &p) {
    p.sum_dt = s.sum_dt;
    for (int k = 0; k < 3; k++) { p.delta_p[k] = s.delta_p[k]; p.delta_v[k] = s.delta_v[k]; p.linearized_ba[k] = s.linearized_ba[k]; p.linearized_bg[k] = s.linearized_bg[k]; }
    p.delta_q[0] = s.delta_q.x; p.delta_q[1] = s.delta_q.y; p.delta_q[2] = s.delta_q.z; p.delta_q[3] = s.delta_q.w;
    for (int j = 0; j < 4; j++) { for (int k = 0; k < 3; k++) p.delta_epsilon[3 * j + k] = s.delta_epsilon[j][k]; p.linearized_rho[j] = s.linearized_rho[j]; }
    for (int r = 0; r < 31; r++) for (int c = 0; c < 31; c++) { p.jacobian[c * 31 + r] = s.jacobian(r, c); p.covariance[c * 31 + r] = s.covariance(r, c); }
}

ImuPreintState to_imu_state(const CerbIMUPreint &p) {
    ImuPreintState s;
    s.sum_dt = p.sum_dt; s.delta_p = V3(p.delta_p); s.delta_v = V3(p.delta_v);
    s.delta_q = Quat(p.delta_q[3], p.delta_q[0], p.delta_q[1], p.delta_q[2]);
    s.linearized_ba = V3(p.linearized_ba); s.linearized_bg = V3(p.linearized_bg);
    for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) { s.jacobian(r, c) = p.jacobian[c * 15 + r]; s.covariance(r, c) = p.covariance[c * 15 + r]; }
    return s;
}
void from_imu_state(const ImuPreintState &s, CerbIMUPreint &p) {
    p.sum_dt = s.sum_dt;
    for (int k = 0; k < 3; k++) { p.delta_p[k] = s.delta_p[k]; p.delta_v[k] = s.delta_v[k]; p.linearized_ba[k] = s.linearized_ba[k]; p.linearized_bg[k] = s.linearized_bg[k]; }
    p.delta_q[0] = s.delta_q.x; p.delta_q[1] = s.delta_q.y; p.delta_q[2] = s.delta_q.z; p.delta_q[3] = s.delta_q.w;
    for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) { p.jacobian[c * 15 + r] = s.jacobian(r, c); p.covariance[c * 15 + r] = s.covariance(r, c); }
}

double *block_ptr(CerbWindowState &st, int kind, int index) {
    switch (kind) {
        case CERB_BLOCK_POSE: return st.para_Pose[index];
        case CERB_BLOCK_SPEEDBIAS: return st.para_SpeedBias[index];
        case CERB_BLOCK_LEGBIAS: return st.para_LegBias[index];
        case CERB_BLOCK_EX_POSE: return st.para_Ex_Pose[index];
        default: return st.para_Td;
    }
}
int block_size(int kind) { return kind == CERB_BLOCK_POSE || kind == CERB_BLOCK_EX_POSE ? 7 : (kind == CERB_BLOCK_SPEEDBIAS ? 9 : (kind == CERB_BLOCK_LEGBIAS ? 4 : 1)); }

void prior_to_info(const CerbPrior &pr, MargInfoLite &info) {
    info.n = pr.n; info.m = 0;
    info.linearized_jacobians = Mat(pr.n, pr.n);
    for (int r = 0; r < pr.n; r++) for (int c = 0; c < pr.n; c++) info.linearized_jacobians(r, c) = pr.linearized_jacobians[(size_t)c * pr.n + r];
    info.linearized_residuals.assign(pr.linearized_residuals, pr.linearized_residuals + pr.n);
    for (int b = 0; b < pr.num_blocks; b++) {
        int sz = block_size(pr.block_kind[b]);
        info.keep_block_size.push_back(sz); info.keep_block_idx.push_back(pr.block_col[b]);
        info.keep_block_data.push_back(std::vector<double>(pr.block_x0[b], pr.block_x0[b] + sz));
    }
}

ProjConst make_proj_const(const CerbObservation &o0, const double *ptj, const double *velj, double tdj) {
    ProjConst c;
    c.pts_i = V3(o0.point[0], o0.point[1], 1.0); c.velocity_i = V3(o0.velocity[0], o0.velocity[1], 0.0); c.td_i = o0.cur_td;
    c.pts_j = V3(ptj[0], ptj[1], 1.0); c.velocity_j = V3(velj[0], velj[1], 0.0); c.td_j = tdj;
    return c;
}

struct WindowProblem {
    Problem problem;
    std::vector<LegPreintState> pre;
    std::vector<ImuPreintState> imu_pre;
    MargInfoLite prior_info;
    FactorGlobals fg;
};

// estimator.cpp:1059-1216
void build_problem(const CerbSolverConfig &cfg, const CerbWindowDesc &d, CerbWindowState &st, WindowProblem &wp) {
    Problem &problem = wp.problem;
    wp.fg.G = V3(cfg.g); wp.fg.visual_sqrt_info = cfg.visual_sqrt_info;
    const bool use_leg = d.preint != nullptr;
    for (int i = 0; i < CERB_NUM_FRAMES; i++) {                       // :1065-1083
        problem.AddParameterBlock(st.para_Pose[i], 7, true, 1);
        problem.AddParameterBlock(st.para_SpeedBias[i], 9, false, 1);
        if (!use_leg) continue;                                        // USE_LEG == 0: no leg-bias blocks at all (:1071)
        problem.AddParameterBlock(st.para_LegBias[i], 4, false, 1);
        if (!cfg.optimize_leg_bias) problem.SetParameterBlockConstant(st.para_LegBias[i]);
    }
    for (int i = 0; i < 2; i++) {                                      // :1087-1101
        problem.AddParameterBlock(st.para_Ex_Pose[i], 7, true, 1);
        if (!d.extrinsic_open) problem.SetParameterBlockConstant(st.para_Ex_Pose[i]);
    }
    problem.AddParameterBlock(st.para_Td, 1, false, 1);                 // :1102-1105
    if (!d.td_open) problem.SetParameterBlockConstant(st.para_Td);
    for (int f = 0; f < d.n_features; f++) problem.AddParameterBlock(&st.para_Feature[f], 1, false, 0);

    if (d.prior.valid) {                                               // :1107-1113
        prior_to_info(d.prior, wp.prior_info);
        std::vector<double *> blocks;
        for (int b = 0; b < d.prior.num_blocks; b++) blocks.push_back(block_ptr(st, d.prior.block_kind[b], d.prior.block_index[b]));
        problem.AddResidualBlock(std::make_shared<MarginalizationFactor>(&wp.prior_info), false, blocks);
    }
    wp.pre.reserve(CERB_WINDOW_SIZE); wp.imu_pre.reserve(CERB_WINDOW_SIZE);
    for (int i = 0; i < CERB_WINDOW_SIZE && !use_leg; i++) {           // :1160-1171  (USE_IMU only)
        wp.imu_pre.push_back(to_imu_state(d.imu_preint[i]));
        if (wp.imu_pre.back().sum_dt > 10.0) continue;
        int j = i + 1;
        problem.AddResidualBlock(std::make_shared<IMUFactor>(&wp.imu_pre[i], wp.fg), false, {st.para_Pose[i], st.para_SpeedBias[i], st.para_Pose[j], st.para_SpeedBias[j]});
    }
    for (int i = 0; i < CERB_WINDOW_SIZE && use_leg; i++) {            // :1114-1159
        wp.pre.push_back(to_state(d.preint[i]));
        if (wp.pre.back().sum_dt > 10.0) continue;
        int j = i + 1;
        problem.AddResidualBlock(std::make_shared<IMULegFactor>(&wp.pre[i], wp.fg), false,
                                 {st.para_Pose[i], st.para_SpeedBias[i], st.para_LegBias[i], st.para_Pose[j], st.para_SpeedBias[j], st.para_LegBias[j]});
    }
    for (int f = 0; f < d.n_features; f++) {                           // :1173-1216
        const CerbFeature &ft = d.features[f];
        const CerbObservation &o0 = d.obs[ft.obs_offset];
        int imu_i = ft.start_frame;
        for (int k = 0; k < ft.n_obs; k++) {
            int imu_j = imu_i + k;
            const CerbObservation &o = d.obs[ft.obs_offset + k];
            if (imu_i != imu_j) {
                ProjConst c = make_proj_const(o0, o.point, o.velocity, o.cur_td);
                problem.AddResidualBlock(std::make_shared<ProjTwoFrameOneCam>(c, wp.fg.visual_sqrt_info), true,
                                         {st.para_Pose[imu_i], st.para_Pose[imu_j], st.para_Ex_Pose[0], &st.para_Feature[f], st.para_Td});
            }
            if (o.is_stereo) {
                ProjConst c = make_proj_const(o0, o.pointRight, o.velocityRight, o.cur_td);
                if (imu_i != imu_j)
                    problem.AddResidualBlock(std::make_shared<ProjTwoFrameTwoCam>(c, wp.fg.visual_sqrt_info), true,
                                             {st.para_Pose[imu_i], st.para_Pose[imu_j], st.para_Ex_Pose[0], st.para_Ex_Pose[1], &st.para_Feature[f], st.para_Td});
                else
                    problem.AddResidualBlock(std::make_shared<ProjOneFrameTwoCam>(c, wp.fg.visual_sqrt_info), true,
                                             {st.para_Ex_Pose[0], st.para_Ex_Pose[1], &st.para_Feature[f], st.para_Td});
            }
        }
    }
}

SolverOptions make_options(const CerbSolverConfig &cfg) {
    SolverOptions o;
    o.max_num_iterations = cfg.max_num_iterations;
    o.initial_trust_region_radius = cfg.initial_trust_region_radius; o.max_trust_region_radius = cfg.max_trust_region_radius;
    o.min_trust_region_radius = cfg.min_trust_region_radius; o.min_relative_decrease = cfg.min_relative_decrease;
    o.function_tolerance = cfg.function_tolerance; o.gradient_tolerance = cfg.gradient_tolerance; o.parameter_tolerance = cfg.parameter_tolerance;
    o.huber_delta = cfg.huber_delta;
    return o;
}

// ---- marginalization: ResidualBlockInfo / MarginalizationInfo (marginalization_factor.cpp) -------
struct RBInfo {
    std::shared_ptr<CostFunction> cost; bool huber; std::vector<double *> params; std::vector<int> drop_set;
    std::vector<double> residuals; std::vector<std::vector<double>> jacobians;   // row-major nres x size
    void Evaluate(double huber_delta) {                                     // :12-78
        int nr = cost->num_residuals, np = (int)params.size();
        residuals.assign(nr, 0.0); jacobians.resize(np);
        std::vector<double *> raw(np);
        for (int i = 0; i < np; i++) { jacobians[i].assign((size_t)nr * cost->block_sizes[i], 0.0); raw[i] = jacobians[i].data(); }
        cost->Evaluate(params.data(), residuals.data(), raw.data());
        if (huber) {
            double sq_norm = 0; for (double v : residuals) sq_norm += v * v;
            double rho[3]; huber_loss(huber_delta, sq_norm, rho);
            double sqrt_rho1_ = std::sqrt(rho[1]), residual_scaling_, alpha_sq_norm_;
            if (sq_norm == 0.0 || rho[2] <= 0.0) { residual_scaling_ = sqrt_rho1_; alpha_sq_norm_ = 0.0; }
            else { const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1]; const double alpha = 1.0 - std::sqrt(D); residual_scaling_ = sqrt_rho1_ / (1 - alpha); alpha_sq_norm_ = alpha / sq_norm; }
            for (int i = 0; i < np; i++) {
                int sz = cost->block_sizes[i];
                for (int c = 0; c < sz; c++) {
                    double rtj = 0; for (int r = 0; r < nr; r++) rtj += residuals[r] * jacobians[i][r * sz + c];
                    for (int r = 0; r < nr; r++) jacobians[i][r * sz + c] = sqrt_rho1_ * (jacobians[i][r * sz + c] - alpha_sq_norm_ * residuals[r] * rtj);
                }
            }
            for (double &v : residuals) v *= residual_scaling_;
        }
    }
};

// marginalization_factor.cpp:281-305: eigen pseudo-inverse of Amm, Schur complement, factoring into (linearized_jacobians, linearized_residuals)
static void marg_schur(const Mat &A, const std::vector<double> &b, int m, int n, double eps, Mat &linearized_jacobians, std::vector<double> &linearized_residuals) {
    Mat Amm(m, m);
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Amm(i, j) = 0.5 * (A(i, j) + A(j, i));
    std::vector<double> ev; Mat V;
    sym_eig(Amm, ev, V);
    Mat Amm_inv(m, m);
    for (int k = 0; k < m; k++) { if (!(ev[k] > eps)) continue; double inv = 1.0 / ev[k];
        for (int i = 0; i < m; i++) { double vi = V(i, k) * inv; for (int j = 0; j < m; j++) Amm_inv(i, j) += vi * V(j, k); } }
    // A = Arr - Arm * Amm_inv * Amr ; b = brr - Arm * Amm_inv * bmm
    Mat Arm(n, m), Amr(m, n);
    for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) { Arm(i, j) = A(m + i, j); Amr(j, i) = A(j, m + i); }
    Mat T = matmul(Arm, Amm_inv);
    Mat TA = matmul(T, Amr);
    Mat Ar(n, n); std::vector<double> br(n);
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) Ar(i, j) = A(m + i, m + j) - TA(i, j);
        double s = b[m + i]; for (int j = 0; j < m; j++) s -= T(i, j) * b[j]; br[i] = s; }
    std::vector<double> ev2; Mat V2;
    Mat Asym(n, n); for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Asym(i, j) = Ar(i, j);   // SelfAdjointEigenSolver reads the lower triangle
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) Asym(i, j) = Asym(j, i);
    sym_eig(Asym, ev2, V2);
    linearized_jacobians = Mat(n, n); linearized_residuals.assign(n, 0.0);
    for (int k = 0; k < n; k++) {
        double S = ev2[k] > eps ? ev2[k] : 0.0, Sinv = ev2[k] > eps ? 1.0 / ev2[k] : 0.0;
        double ss = std::sqrt(S), sis = std::sqrt(Sinv);
        double vb = 0; for (int i = 0; i < n; i++) vb += V2(i, k) * br[i];
        for (int j = 0; j < n; j++) linearized_jacobians(k, j) = ss * V2(j, k);
        linearized_residuals[k] = sis * vb;
    }
}

struct MargInfo {
    std::vector<RBInfo> factors;
    std::vector<double *> order;                         // insertion order of parameter blocks (replaces unordered_map order)
    std::map<double *, int> size, idx; std::map<double *, bool> dropped;
    std::map<double *, std::vector<double>> data;
    int m = 0, n = 0; bool valid = true;
    Mat linearized_jacobians; std::vector<double> linearized_residuals;
    const double eps = 1e-8;
    void add(RBInfo &&f) {                                // addResidualBlockInfo :98-117
        for (size_t i = 0; i < f.params.size(); i++) { if (!size.count(f.params[i])) order.push_back(f.params[i]); size[f.params[i]] = f.cost->block_sizes[i]; }
        for (int dset : f.drop_set) dropped[f.params[dset]] = true;
        factors.push_back(std::move(f));
    }
    static int localSize(int s) { return s == 7 ? 6 : s; }
    void preMarginalize(double huber_delta) {             // :119-138
        for (auto &f : factors) {
            f.Evaluate(huber_delta);
            for (size_t i = 0; i < f.params.size(); i++)
                if (!data.count(f.params[i])) data[f.params[i]] = std::vector<double>(f.params[i], f.params[i] + f.cost->block_sizes[i]);
        }
    }
    void marginalize() {                                  // :183-305
        int pos = 0;
        for (double *p : order) if (dropped.count(p)) { idx[p] = pos; pos += localSize(size[p]); }
        m = pos;
        for (double *p : order) if (!dropped.count(p)) { idx[p] = pos; pos += localSize(size[p]); }
        n = pos - m;
        if (m == 0) { valid = false; return; }
        Mat A(pos, pos); std::vector<double> b(pos, 0.0);
        for (auto &f : factors) {                          // ThreadsConstructA :150-181 (summed in one thread)
            int nr = f.cost->num_residuals;
            for (size_t i = 0; i < f.params.size(); i++) {
                int idx_i = idx[f.params[i]], gi = f.cost->block_sizes[i], size_i = localSize(gi);
                for (size_t j = i; j < f.params.size(); j++) {
                    int idx_j = idx[f.params[j]], gj = f.cost->block_sizes[j], size_j = localSize(gj);
                    for (int a = 0; a < size_i; a++) for (int c = 0; c < size_j; c++) {
                        double s = 0; for (int r = 0; r < nr; r++) s += f.jacobians[i][r * gi + a] * f.jacobians[j][r * gj + c];
                        A(idx_i + a, idx_j + c) += s;
                        if (i != j) A(idx_j + c, idx_i + a) = A(idx_i + a, idx_j + c);
                    }
                }
                for (int a = 0; a < size_i; a++) { double s = 0; for (int r = 0; r < nr; r++) s += f.jacobians[i][r * gi + a] * f.residuals[r]; b[idx_i + a] += s; }
            }
        }
        marg_schur(A, b, m, n, eps, linearized_jacobians, linearized_residuals);
    }
};

struct OwnedPrior { CerbPrior p; std::vector<double> J, r; };

}  // namespace

extern "C" {

int oracle_solve_window(const CerbSolverConfig *cfg, const CerbWindowDesc *desc, CerbWindowState *state, CerbSolveReport *report,
                        double *gradient0, double *jtj_diag0, int n_alloc) {
    WindowProblem wp;
    build_problem(*cfg, *desc, *state, wp);
    SolverSummary s;
    Solve(make_options(*cfg), wp.problem, s);
    if (report) { report->iterations = s.iterations; report->num_successful_steps = s.num_successful_steps; report->termination = s.termination;
                  report->status = std::isfinite(s.final_cost) ? CERB_OK : CERB_ERR_NON_FINITE; report->initial_cost = s.initial_cost; report->final_cost = s.final_cost; }
    if (gradient0 || jtj_diag0) {
        // ABI order: [pose0..10 (66) | ex0, ex1 (12) | speedbias0..10 (99) | legbias0..10 (44) | td | features]; constant blocks -> 0
        int total = 222 + desc->n_features;
        if (n_alloc < total) return CERB_ERR_BAD_ARGUMENT;
        for (int k = 0; k < total; k++) { if (gradient0) gradient0[k] = 0; if (jtj_diag0) jtj_diag0[k] = 0; }
        auto put = [&](double *ptr, int off) {
            if (!wp.problem.index.count(ptr)) return;
            const Problem::PB &b = wp.problem.pbs[wp.problem.index.at(ptr)];
            if (b.constant || s.gradient0.empty()) return;
            for (int k = 0; k < b.local; k++) { if (gradient0) gradient0[off + k] = s.gradient0[b.col + k]; if (jtj_diag0) jtj_diag0[off + k] = s.jtj_diag0[b.col + k]; }
        };
        for (int i = 0; i < 11; i++) { put(state->para_Pose[i], 6 * i); put(state->para_SpeedBias[i], 78 + 9 * i); put(state->para_LegBias[i], 177 + 4 * i); }
        put(state->para_Ex_Pose[0], 66); put(state->para_Ex_Pose[1], 72); put(state->para_Td, 221);
        for (int f = 0; f < desc->n_features; f++) put(&state->para_Feature[f], 222 + f);
    }
    return CERB_OK;
}

// CPU baseline: one window per thread, each solve single-threaded like the reference (estimator.cpp:1224)
int oracle_solve_batch(const CerbSolverConfig *cfg, int n, const CerbWindowDesc *descs, CerbWindowState *states, CerbSolveReport *reports, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    auto work = [&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; oracle_solve_window(cfg, &descs[i], &states[i], reports ? &reports[i] : nullptr, nullptr, nullptr, 0); } };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    return CERB_OK;
}

int oracle_eval_projection(int kind, int n, double sqrt_info, const double *pose_i, const double *pose_j, const double *ex0, const double *ex1,
                           const double *inv_dep, const double *td, const double *pts_i, const double *pts_j, const double *vel_i,
                           const double *vel_j, const double *td_i, const double *td_j, double *residuals, double *jacobians) {
    for (int k = 0; k < n; k++) {
        ProjConst c;
        c.pts_i = V3(pts_i + 3 * k); c.pts_j = V3(pts_j + 3 * k);
        c.velocity_i = V3(vel_i[2 * k], vel_i[2 * k + 1], 0); c.velocity_j = V3(vel_j[2 * k], vel_j[2 * k + 1], 0);
        c.td_i = td_i[k]; c.td_j = td_j[k];
        double r[2]; double *J[6]; const double *p[6];
        if (kind == CERB_PROJ_TWO_FRAME_ONE_CAM) {
            ProjTwoFrameOneCam f(c, sqrt_info);
            double *base = jacobians ? jacobians + (size_t)k * 46 : nullptr;
            p[0] = pose_i + 7 * k; p[1] = pose_j + 7 * k; p[2] = ex0 + 7 * k; p[3] = inv_dep + k; p[4] = td + k;
            if (base) { J[0] = base; J[1] = base + 14; J[2] = base + 28; J[3] = base + 42; J[4] = base + 44; }
            f.Evaluate(p, r, base ? J : nullptr);
        } else if (kind == CERB_PROJ_TWO_FRAME_TWO_CAM) {
            ProjTwoFrameTwoCam f(c, sqrt_info);
            double *base = jacobians ? jacobians + (size_t)k * 60 : nullptr;
            p[0] = pose_i + 7 * k; p[1] = pose_j + 7 * k; p[2] = ex0 + 7 * k; p[3] = ex1 + 7 * k; p[4] = inv_dep + k; p[5] = td + k;
            if (base) { J[0] = base; J[1] = base + 14; J[2] = base + 28; J[3] = base + 42; J[4] = base + 56; J[5] = base + 58; }
            f.Evaluate(p, r, base ? J : nullptr);
        } else {
            ProjOneFrameTwoCam f(c, sqrt_info);
            double *base = jacobians ? jacobians + (size_t)k * 32 : nullptr;
            p[0] = ex0 + 7 * k; p[1] = ex1 + 7 * k; p[2] = inv_dep + k; p[3] = td + k;
            if (base) { J[0] = base; J[1] = base + 14; J[2] = base + 28; J[3] = base + 30; }
            f.Evaluate(p, r, base ? J : nullptr);
        }
        if (residuals) { residuals[2 * k] = r[0]; residuals[2 * k + 1] = r[1]; }
    }
    return CERB_OK;
}

int oracle_eval_imu_leg(int n, const double g[3], const CerbIMULegPreint *preint, const double *params, double *residuals, double *jacobians, double *sqrt_info) {
    FactorGlobals fg; fg.G = V3(g);
    for (int k = 0; k < n; k++) {
        LegPreintState s = to_state(preint[k]);
        IMULegFactor f(&s, fg);
        const double *q = params + (size_t)k * 40;
        const double *p[6] = {q, q + 7, q + 16, q + 20, q + 27, q + 36};
        double r[31]; double *J[6];
        double *base = jacobians ? jacobians + (size_t)k * 31 * 40 : nullptr;
        if (base) { J[0] = base; J[1] = base + 31 * 7; J[2] = base + 31 * 16; J[3] = base + 31 * 20; J[4] = base + 31 * 27; J[5] = base + 31 * 36; }
        if (!f.Evaluate(p, r, base ? J : nullptr)) return CERB_ERR_NON_FINITE;
        if (residuals) for (int i = 0; i < 31; i++) residuals[(size_t)k * 31 + i] = r[i];
        if (sqrt_info) { Mat si; imu_leg_sqrt_info(s.covariance, si); for (int i = 0; i < 31 * 31; i++) sqrt_info[(size_t)k * 961 + i] = si.d[i]; }
    }
    return CERB_OK;
}

int oracle_eval_imu(int n, const double g[3], const CerbIMUPreint *preint, const double *params, double *residuals, double *jacobians, double *sqrt_info) {
    FactorGlobals fg; fg.G = V3(g);
    for (int k = 0; k < n; k++) {
        ImuPreintState s = to_imu_state(preint[k]);
        IMUFactor f(&s, fg);
        const double *q = params + (size_t)k * 32;
        const double *p[4] = {q, q + 7, q + 16, q + 23};
        double r[15]; double *J[4];
        double *base = jacobians ? jacobians + (size_t)k * 15 * 32 : nullptr;
        if (base) { J[0] = base; J[1] = base + 15 * 7; J[2] = base + 15 * 16; J[3] = base + 15 * 23; }
        if (!f.Evaluate(p, r, base ? J : nullptr)) return CERB_ERR_NON_FINITE;
        if (residuals) for (int i = 0; i < 15; i++) residuals[(size_t)k * 15 + i] = r[i];
        if (sqrt_info) { Mat si; imu_leg_sqrt_info(s.covariance, si); for (int i = 0; i < 225; i++) sqrt_info[(size_t)k * 225 + i] = si.d[i]; }
    }
    return CERB_OK;
}

int oracle_eval_prior(const CerbPrior *prior, const CerbWindowState *state, double *residuals, double *jacobians) {
    MargInfoLite info; prior_to_info(*prior, info);
    MarginalizationFactor f(&info);
    CerbWindowState st = *state;
    std::vector<const double *> p; std::vector<double *> J;
    size_t off = 0;
    for (int b = 0; b < prior->num_blocks; b++) {
        p.push_back(block_ptr(st, prior->block_kind[b], prior->block_index[b]));
        J.push_back(jacobians ? jacobians + off : nullptr);
        off += (size_t)prior->n * block_size(prior->block_kind[b]);
    }
    std::vector<double> r(prior->n);
    f.Evaluate(p.data(), r.data(), jacobians ? J.data() : nullptr);
    if (residuals) for (int i = 0; i < prior->n; i++) residuals[i] = r[i];
    return CERB_OK;
}

static PreintGlobals to_globals(const CerbPreintConfig &c) {
    PreintGlobals g;
    g.ACC_N = c.acc_n; g.ACC_N_Z = c.acc_n_z; g.GYR_N = c.gyr_n; g.ACC_W = c.acc_w; g.GYR_W = c.gyr_w;
    g.PHI_N = c.phi_n; g.DPHI_N = c.dphi_n; g.RHO_C_N = c.rho_c_n; g.RHO_NC_N = c.rho_nc_n;
    g.V_N_MIN_XY = c.v_n_min_xy; g.V_N_MIN_Z = c.v_n_min_z; g.V_N_MIN = c.v_n_min; g.V_N_MAX = c.v_n_max;
    g.V_N_FORCE_THRES_RATIO = c.v_n_force_thres_ratio; g.V_N_TERM1_STEEP = c.v_n_term1_steep;
    g.V_N_TERM2_VAR_RESCALE = c.v_n_term2_var_rescale; g.V_N_TERM3_DISTANCE_RESCALE = c.v_n_term3_distance_rescale;
    g.CONTACT_SENSOR_TYPE = c.contact_sensor_type;
    for (int l = 0; l < 4; l++) for (int k = 0; k < 4; k++) g.rho_fix[l][k] = c.rho_fix[l][k];
    g.p_br = V3(c.p_br);
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) g.R_br(r, k) = c.R_br[3 * r + k];
    return g;
}

int oracle_preintegrate(const CerbPreintConfig *cfg, int n, const CerbPreintJob *jobs, CerbIMULegPreint *out) {
    PreintGlobals g = to_globals(*cfg);
    for (int k = 0; k < n; k++) {
        const CerbPreintJob &j = jobs[k];
        LegPreintegrator pi(g, V3(j.acc_0), V3(j.gyr_0), j.phi_0, j.dphi_0, j.c_0, V3(j.linearized_ba), V3(j.linearized_bg), j.linearized_rho);
        for (int s = 0; s < j.n_samples; s++) { const CerbIMULegSample &m = j.samples[s]; pi.push_back(m.dt, V3(m.acc), V3(m.gyr), m.phi, m.dphi, m.c); }
        from_state(pi, out[k]);
    }
    return CERB_OK;
}

int oracle_preintegrate_imu(const CerbPreintConfig *cfg, int n, const CerbPreintJob *jobs, CerbIMUPreint *out) {
    PreintGlobals g = to_globals(*cfg);
    for (int k = 0; k < n; k++) {
        const CerbPreintJob &j = jobs[k];
        ImuPreintegrator pi(g, V3(j.acc_0), V3(j.gyr_0), V3(j.linearized_ba), V3(j.linearized_bg));
        for (int s = 0; s < j.n_samples; s++) pi.push_back(j.samples[s].dt, V3(j.samples[s].acc), V3(j.samples[s].gyr));
        from_imu_state(pi, out[k]);
    }
    return CERB_OK;
}

int oracle_a1_kinematics(int n, const double *q, const double *rho_opt, const double *rho_fix, double *fk, double *jac, double *dfk_drho, double *dJ_dq, double *dJ_drho) {
    for (int k = 0; k < n; k++) {
        if (fk) a1_fk(q + 3 * k, rho_opt[k], rho_fix + 4 * k, fk + 3 * k);
        if (jac) a1_jac(q + 3 * k, rho_opt[k], rho_fix + 4 * k, jac + 9 * k);
        if (dfk_drho) a1_dfk_drho(q + 3 * k, rho_opt[k], rho_fix + 4 * k, dfk_drho + 3 * k);
        if (dJ_dq) a1_dJ_dq(q + 3 * k, rho_opt[k], rho_fix + 4 * k, dJ_dq + 27 * k);
        if (dJ_drho) a1_dJ_drho(q + 3 * k, rho_opt[k], rho_fix + 4 * k, dJ_drho + 9 * k);
    }
    return CERB_OK;
}

// Marginalization glue of optimization(): margin_old != 0 -> estimator.cpp:1248-1376, else :1377-1455.
// Writes the new prior (already address-shifted for the NEXT window) into *out; J_out (>= 96*96) and
// r_out (>= 96) receive linearized_jacobians (column-major n x n) / linearized_residuals.
// Returns CERB_OK; out->valid == 0 if nothing was marginalized.
int oracle_marginalize(const CerbSolverConfig *cfg, const CerbWindowDesc *desc, const CerbWindowState *state_in, int margin_old,
                       CerbPrior *out, double *J_out, double *r_out) {
    CerbWindowState st = *state_in;
    std::vector<double> feat(state_in->para_Feature, state_in->para_Feature + desc->n_features);
    st.para_Feature = feat.data();
    FactorGlobals fg; fg.G = V3(cfg->g); fg.visual_sqrt_info = cfg->visual_sqrt_info;
    MargInfo mi;
    MargInfoLite last; std::vector<LegPreintState> pre; std::vector<ImuPreintState> ipre;
    std::memset(out, 0, sizeof(*out));
    if (margin_old) {
        if (desc->prior.valid) {
            prior_to_info(desc->prior, last);
            RBInfo f; f.cost = std::make_shared<MarginalizationFactor>(&last); f.huber = false;
            for (int b = 0; b < desc->prior.num_blocks; b++) {
                f.params.push_back(block_ptr(st, desc->prior.block_kind[b], desc->prior.block_index[b]));
                int kind = desc->prior.block_kind[b], index = desc->prior.block_index[b];
                if (index == 0 && (kind == CERB_BLOCK_POSE || kind == CERB_BLOCK_SPEEDBIAS || kind == CERB_BLOCK_LEGBIAS)) f.drop_set.push_back(b);
            }
            mi.add(std::move(f));
        }
        if (desc->preint) {                                          // USE_LEG: estimator.cpp:1271-1285
            pre.push_back(to_state(desc->preint[0]));
            if (pre[0].sum_dt < 10.0) {
                RBInfo f; f.cost = std::make_shared<IMULegFactor>(&pre[0], fg); f.huber = false;
                f.params = {st.para_Pose[0], st.para_SpeedBias[0], st.para_LegBias[0], st.para_Pose[1], st.para_SpeedBias[1], st.para_LegBias[1]};
                f.drop_set = {0, 1, 2};
                mi.add(std::move(f));
            }
        } else {                                                     // USE_IMU only: IMUFactor <15,7,9,7,9>, drop pose0 / speedbias0 (estimator.cpp:1287-1297)
            ipre.push_back(to_imu_state(desc->imu_preint[0]));
            if (ipre[0].sum_dt < 10.0) {
                RBInfo f; f.cost = std::make_shared<IMUFactor>(&ipre[0], fg); f.huber = false;
                f.params = {st.para_Pose[0], st.para_SpeedBias[0], st.para_Pose[1], st.para_SpeedBias[1]};
                f.drop_set = {0, 1};
                mi.add(std::move(f));
            }
        }
        for (int fi = 0; fi < desc->n_features; fi++) {
            const CerbFeature &ft = desc->features[fi];
            if (ft.start_frame != 0) continue;
            const CerbObservation &o0 = desc->obs[ft.obs_offset];
            for (int k = 0; k < ft.n_obs; k++) {
                int imu_j = k;
                const CerbObservation &o = desc->obs[ft.obs_offset + k];
                if (imu_j != 0) {
                    RBInfo f; f.cost = std::make_shared<ProjTwoFrameOneCam>(make_proj_const(o0, o.point, o.velocity, o.cur_td), fg.visual_sqrt_info); f.huber = true;
                    f.params = {st.para_Pose[0], st.para_Pose[imu_j], st.para_Ex_Pose[0], &st.para_Feature[fi], st.para_Td}; f.drop_set = {0, 3};
                    mi.add(std::move(f));
                }
                if (o.is_stereo) {
                    ProjConst c = make_proj_const(o0, o.pointRight, o.velocityRight, o.cur_td);
                    RBInfo f; f.huber = true;
                    if (imu_j != 0) { f.cost = std::make_shared<ProjTwoFrameTwoCam>(c, fg.visual_sqrt_info);
                        f.params = {st.para_Pose[0], st.para_Pose[imu_j], st.para_Ex_Pose[0], st.para_Ex_Pose[1], &st.para_Feature[fi], st.para_Td}; f.drop_set = {0, 4}; }
                    else { f.cost = std::make_shared<ProjOneFrameTwoCam>(c, fg.visual_sqrt_info);
                        f.params = {st.para_Ex_Pose[0], st.para_Ex_Pose[1], &st.para_Feature[fi], st.para_Td}; f.drop_set = {2}; }
                    mi.add(std::move(f));
                }
            }
        }
    } else {
        bool has = false;
        if (desc->prior.valid) for (int b = 0; b < desc->prior.num_blocks; b++) if (desc->prior.block_kind[b] == CERB_BLOCK_POSE && desc->prior.block_index[b] == CERB_WINDOW_SIZE - 1) has = true;
        if (!has) { *out = desc->prior; return CERB_OK; }   // prior carried over unchanged (pointers alias the input)
        prior_to_info(desc->prior, last);
        RBInfo f; f.cost = std::make_shared<MarginalizationFactor>(&last); f.huber = false;
        for (int b = 0; b < desc->prior.num_blocks; b++) {
            f.params.push_back(block_ptr(st, desc->prior.block_kind[b], desc->prior.block_index[b]));
            if (desc->prior.block_kind[b] == CERB_BLOCK_POSE && desc->prior.block_index[b] == CERB_WINDOW_SIZE - 1) f.drop_set.push_back(b);
        }
        mi.add(std::move(f));
    }
    mi.preMarginalize(cfg->huber_delta);
    mi.marginalize();
    if (!mi.valid) { out->valid = 0; return CERB_OK; }
    // getParameterBlocks + addr_shift
    out->valid = 1; out->n = mi.n; out->num_blocks = 0;
    for (double *p : mi.order) {
        if (mi.dropped.count(p)) continue;
        int kind = -1, index = -1;
        for (int i = 0; i < CERB_NUM_FRAMES; i++) {
            if (p == st.para_Pose[i]) { kind = CERB_BLOCK_POSE; index = i; }
            if (p == st.para_SpeedBias[i]) { kind = CERB_BLOCK_SPEEDBIAS; index = i; }
            if (p == st.para_LegBias[i]) { kind = CERB_BLOCK_LEGBIAS; index = i; }
        }
        for (int i = 0; i < 2; i++) if (p == st.para_Ex_Pose[i]) { kind = CERB_BLOCK_EX_POSE; index = i; }
        if (p == st.para_Td) { kind = CERB_BLOCK_TD; index = 0; }
        if (kind < 0) continue;   // cannot happen: features are always dropped
        if (kind <= CERB_BLOCK_LEGBIAS) {
            if (margin_old) index -= 1;                                    // :1359-1364
            else if (index == CERB_WINDOW_SIZE) index -= 1;                // :1421-1431
        }
        int b = out->num_blocks++;
        if (b >= CERB_MAX_PRIOR_BLOCKS) return CERB_ERR_BAD_ARGUMENT;
        out->block_kind[b] = kind; out->block_index[b] = index; out->block_col[b] = mi.idx[p] - mi.m;
        const std::vector<double> &dv = mi.data[p];
        for (size_t k = 0; k < dv.size(); k++) out->block_x0[b][k] = dv[k];
    }
    for (int r = 0; r < mi.n; r++) { r_out[r] = mi.linearized_residuals[r]; for (int c = 0; c < mi.n; c++) J_out[(size_t)c * mi.n + r] = mi.linearized_jacobians(r, c); }
    out->linearized_jacobians = J_out; out->linearized_residuals = r_out;
    return CERB_OK;
}

// Estimator::double2vector gauge re-anchoring, estimator.cpp:903-957
void oracle_double2vector(const CerbWindowState *before, const CerbWindowState *after, double *Ps, double *Rs, double *Vs) {
    auto quat_of = [](const double *p) { return Quat(p[6], p[3], p[4], p[5]); };
    M3 Rs0 = toR(quat_of(before->para_Pose[0]));
    V3 origin_R0 = R2ypr(Rs0);
    V3 origin_P0(before->para_Pose[0]);
    V3 origin_R00 = R2ypr(toR(quat_of(after->para_Pose[0])));
    double y_diff = origin_R0.x - origin_R00.x;
    M3 rot_diff = ypr2R(V3(y_diff, 0, 0));
    if (std::fabs(std::fabs(origin_R0.y) - 90) < 1.0 || std::fabs(std::fabs(origin_R00.y) - 90) < 1.0)
        rot_diff = Rs0 * transpose(toR(quat_of(after->para_Pose[0])));
    for (int i = 0; i < CERB_NUM_FRAMES; i++) {
        M3 R = rot_diff * toR(normalized(quat_of(after->para_Pose[i])));
        V3 P = rot_diff * V3(after->para_Pose[i][0] - after->para_Pose[0][0], after->para_Pose[i][1] - after->para_Pose[0][1], after->para_Pose[i][2] - after->para_Pose[0][2]) + origin_P0;
        V3 V = rot_diff * V3(after->para_SpeedBias[i]);
        for (int k = 0; k < 3; k++) { Ps[3 * i + k] = P[k]; Vs[3 * i + k] = V[k]; for (int c = 0; c < 3; c++) Rs[9 * i + 3 * k + c] = R(k, c); }
    }
}

// ---- per-feature steps either side of the solve ------------------------------------------------------
// Estimator::reprojectionError, estimator.cpp:1729-1739
static double reprojection_error(const M3 &Ri, V3 Pi, const M3 &rici, V3 tici, const M3 &Rj, V3 Pj, const M3 &ricj, V3 ticj, double depth, V3 uvi, V3 uvj) {
    V3 pts_w = Ri * (rici * (depth * uvi) + tici) + Pi;
    V3 pts_cj = transpose(ricj) * (transpose(Rj) * (pts_w - Pj) - ticj);
    double rx = pts_cj.x / pts_cj.z - uvj.x, ry = pts_cj.y / pts_cj.z - uvj.y;
    return std::sqrt(rx * rx + ry * ry);
}
// Estimator::outliersRejection, estimator.cpp:1741-1798: ave_err[f] for every feature of the window (Rs, Ps, ric, tic from the
// para_* arrays; depth = estimated_depth = 1 / para_Feature, feature_manager.cpp:189)
int oracle_outlier_errors(const CerbWindowDesc *d, const CerbWindowState *st, double *ave_err) {
    auto quat_of = [](const double *p) { return Quat(p[6], p[3], p[4], p[5]); };
    M3 ric[2] = {toR(quat_of(st->para_Ex_Pose[0])), toR(quat_of(st->para_Ex_Pose[1]))};
    V3 tic[2] = {V3(st->para_Ex_Pose[0]), V3(st->para_Ex_Pose[1])};
    for (int f = 0; f < d->n_features; f++) {
        const CerbFeature &ft = d->features[f];
        double err = 0; int errCnt = 0;
        int imu_i = ft.start_frame, imu_j = imu_i - 1;
        const CerbObservation &o0 = d->obs[ft.obs_offset];
        V3 pts_i(o0.point[0], o0.point[1], 1.0);
        double depth = 1.0 / st->para_Feature[f];
        M3 Ri = toR(quat_of(st->para_Pose[imu_i])); V3 Pi(st->para_Pose[imu_i]);
        for (int k = 0; k < ft.n_obs; k++) {
            const CerbObservation &o = d->obs[ft.obs_offset + k];
            imu_j++;
            M3 Rj = toR(quat_of(st->para_Pose[imu_j])); V3 Pj(st->para_Pose[imu_j]);
            if (imu_i != imu_j) { err += reprojection_error(Ri, Pi, ric[0], tic[0], Rj, Pj, ric[0], tic[0], depth, pts_i, V3(o.point[0], o.point[1], 1.0)); errCnt++; }
            if (o.is_stereo) { err += reprojection_error(Ri, Pi, ric[0], tic[0], Rj, Pj, ric[1], tic[1], depth, pts_i, V3(o.pointRight[0], o.pointRight[1], 1.0)); errCnt++; }
        }
        ave_err[f] = err / errCnt;
    }
    return 0;
}
// Smallest right singular vector of a 4 x 4 matrix (what design_matrix.jacobiSvd(ComputeFullV).matrixV().rightCols<1>() returns up
// to sign): eigenvector of the smallest eigenvalue of A^T A by cyclic two-sided Jacobi rotations.
static void null_vector4(const double A[4][4], double v[4]) {
    double B[4][4], V[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { B[i][j] = 0; for (int k = 0; k < 4; k++) B[i][j] += A[k][i] * A[k][j]; V[i][j] = i == j; }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0; for (int p = 0; p < 4; p++) for (int q = p + 1; q < 4; q++) off += B[p][q] * B[p][q];
        if (off < 1e-60) break;
        for (int p = 0; p < 3; p++) for (int q = p + 1; q < 4; q++) {
            if (B[p][q] == 0.0) continue;
            double th = (B[q][q] - B[p][p]) / (2 * B[p][q]);
            double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), sn = t * c;
            for (int k = 0; k < 4; k++) { double bp = B[k][p], bq = B[k][q]; B[k][p] = c * bp - sn * bq; B[k][q] = sn * bp + c * bq; }
            for (int k = 0; k < 4; k++) { double bp = B[p][k], bq = B[q][k]; B[p][k] = c * bp - sn * bq; B[q][k] = sn * bp + c * bq; }
            for (int k = 0; k < 4; k++) { double vp = V[k][p], vq = V[k][q]; V[k][p] = c * vp - sn * vq; V[k][q] = sn * vp + c * vq; }
        }
    }
    int best = 0; for (int c = 1; c < 4; c++) if (B[c][c] < B[best][best]) best = c;
    for (int k = 0; k < 4; k++) v[k] = V[k][best];
}
// FeatureManager::triangulatePoint, feature_manager.cpp:198-212 (Pose = [R^T | -R^T t], 3 x 4)
static V3 triangulate_point(const double P0[3][4], const double P1[3][4], double u0x, double u0y, double u1x, double u1y) {
    double A[4][4], v[4];
    for (int c = 0; c < 4; c++) { A[0][c] = u0x * P0[2][c] - P0[0][c]; A[1][c] = u0y * P0[2][c] - P0[1][c]; A[2][c] = u1x * P1[2][c] - P1[0][c]; A[3][c] = u1y * P1[2][c] - P1[1][c]; }
    null_vector4(A, v);
    return V3(v[0] / v[3], v[1] / v[3], v[2] / v[3]);
}
// FeatureManager::triangulate, feature_manager.cpp:302-385: depth[f] for the features with estimated_depth <= 0
// (para_Feature <= 0); the others return their current 1 / para_Feature.  (:387-428 is unreachable: size() > 1 takes :351.)
int oracle_triangulate(const CerbWindowDesc *d, const CerbWindowState *st, double init_depth, double *depth) {
    auto quat_of = [](const double *p) { return Quat(p[6], p[3], p[4], p[5]); };
    M3 ric[2] = {toR(quat_of(st->para_Ex_Pose[0])), toR(quat_of(st->para_Ex_Pose[1]))};
    V3 tic[2] = {V3(st->para_Ex_Pose[0]), V3(st->para_Ex_Pose[1])};
    auto pose34 = [](const M3 &R, V3 t, double P[3][4]) { M3 Rt = transpose(R); V3 m = -(Rt * t); for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) P[r][c] = Rt(r, c); P[r][3] = m[r]; } };
    for (int f = 0; f < d->n_features; f++) {
        const double l = st->para_Feature[f];
        if (l > 0) { depth[f] = 1.0 / l; continue; }
        const CerbFeature &ft = d->features[f];
        const CerbObservation &o0 = d->obs[ft.obs_offset];
        int imu_i = ft.start_frame;
        M3 Ri = toR(quat_of(st->para_Pose[imu_i])); V3 Pi(st->para_Pose[imu_i]);
        double left[3][4], right[3][4];
        pose34(Ri * ric[0], Pi + Ri * tic[0], left);
        V3 p;
        if (o0.is_stereo) { pose34(Ri * ric[1], Pi + Ri * tic[1], right); p = triangulate_point(left, right, o0.point[0], o0.point[1], o0.pointRight[0], o0.pointRight[1]); }
        else if (ft.n_obs > 1) {
            const CerbObservation &o1 = d->obs[ft.obs_offset + 1];
            M3 Rj = toR(quat_of(st->para_Pose[imu_i + 1])); V3 Pj(st->para_Pose[imu_i + 1]);
            pose34(Rj * ric[0], Pj + Rj * tic[0], right);
            p = triangulate_point(left, right, o0.point[0], o0.point[1], o1.point[0], o1.point[1]);
        } else { depth[f] = l; continue; }
        double dz = left[2][0] * p.x + left[2][1] * p.y + left[2][2] * p.z + left[2][3];
        depth[f] = dz > 0 ? dz : init_depth;
    }
    return 0;
}

// FeatureManager::removeBackShiftDepth via Estimator::slideWindowOld, feature_manager.cpp:450-488 + estimator.cpp:1660-1677
int oracle_shift_depth(const CerbWindowDesc *d, const CerbWindowState *st, double init_depth, int *new_start, double *depth, int *keep) {
    auto quat_of = [](const double *p) { return Quat(p[6], p[3], p[4], p[5]); };
    M3 ric0 = toR(quat_of(st->para_Ex_Pose[0])); V3 tic0(st->para_Ex_Pose[0]);
    M3 back_R0 = toR(quat_of(st->para_Pose[0])), Rs0 = toR(quat_of(st->para_Pose[1]));      // Rs[0] after the slide is the old frame 1
    V3 back_P0(st->para_Pose[0]), Ps0(st->para_Pose[1]);
    M3 R0 = back_R0 * ric0, R1 = Rs0 * ric0;
    V3 P0 = back_P0 + back_R0 * tic0, P1 = Ps0 + Rs0 * tic0;
    for (int f = 0; f < d->n_features; f++) {
        const CerbFeature &ft = d->features[f];
        double est = 1.0 / st->para_Feature[f];
        keep[f] = 1; depth[f] = est;
        if (ft.start_frame != 0) { new_start[f] = ft.start_frame - 1; continue; }
        new_start[f] = 0;
        if (ft.n_obs - 1 < 2) { keep[f] = 0; continue; }
        const CerbObservation &o0 = d->obs[ft.obs_offset];
        V3 uv_i(o0.point[0], o0.point[1], 1.0);
        V3 pts_i = uv_i * est;
        V3 w_pts_i = R0 * pts_i + P0;
        V3 pts_j = transpose(R1) * (w_pts_i - P1);
        depth[f] = pts_j.z > 0 ? pts_j.z : init_depth;
    }
    return 0;
}

int oracle_marginalize_schur(int m, int n, const double *A_rowmajor, const double *b_in, double eps, double *lin_J_colmajor, double *lin_r) {
    int pos = m + n;
    Mat A(pos, pos); std::vector<double> b(b_in, b_in + pos), r; Mat J;
    for (int i = 0; i < pos; i++) for (int j = 0; j < pos; j++) A(i, j) = A_rowmajor[(size_t)i * pos + j];
    marg_schur(A, b, m, n, eps, J, r);
    for (int k = 0; k < n; k++) { for (int j = 0; j < n; j++) lin_J_colmajor[k + (size_t)j * n] = J(k, j); lin_r[k] = r[k]; }
    return 0;
}

void oracle_set_eig_mode(int mode) { eig_mode() = mode; }   // 0: tridiagonal QR (the reference's Eigen solver, default), 1: cyclic Jacobi

int oracle_abi_sizes(int *out, int n) {   // struct-size handshake for the ctypes mirror
    int v[] = {(int)sizeof(CerbSolverConfig), (int)sizeof(CerbIMULegPreint), (int)sizeof(CerbObservation), (int)sizeof(CerbFeature), (int)sizeof(CerbPrior),
               (int)sizeof(CerbWindowDesc), (int)sizeof(CerbWindowState), (int)sizeof(CerbSolveReport), (int)sizeof(CerbIMULegSample), (int)sizeof(CerbPreintConfig), (int)sizeof(CerbPreintJob), (int)sizeof(CerbIMUPreint)};
    int k = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < k && i < n; i++) out[i] = v[i];
    return k;
}

}  // extern "C"
