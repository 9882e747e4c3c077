// oracle/ref_factors.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// CPU restatement of the Cerberus factor families behind a ceres::CostFunction-shaped interface.
// PARITY UNPINNED (no golden vectors in the reference); self-consistency is checked by
// tests/test_oracle_jacobians.py (analytic vs central-difference, the procedure of
// projectionTwoFrameOneCamFactor.cpp:152-272).
#pragma once
#include "ref_math.h"
#include <memory>

namespace oracle {

// ceres::CostFunction restated: Evaluate(parameters, residuals, jacobians) with row-major
// jacobians[k] of num_residuals x block_sizes[k]; jacobians / jacobians[k] may be null.
struct CostFunction {
    int num_residuals = 0;
    std::vector<int> block_sizes;
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
};

// Globals the factors read (src/utils/parameters.h): G and the static sqrt_info of the
// projection factors (estimator.cpp:124-126).
struct FactorGlobals {
    V3 G{0, 0, 9.805};
    double visual_sqrt_info = 460.0 / 1.5;
};

struct ProjConst {
    V3 pts_i, pts_j;          // z = 1
    V3 velocity_i, velocity_j; // z = 0 (ctor of the reference zeroes it)
    double td_i = 0, td_j = 0;
};

// projectionTwoFrameOneCamFactor.cpp:43-150   SizedCostFunction<2,7,7,7,1,1>
struct ProjTwoFrameOneCam : CostFunction {
    ProjConst c; double sqrt_info;
    ProjTwoFrameOneCam(const ProjConst &c_, double si) : c(c_), sqrt_info(si) { num_residuals = 2; block_sizes = {7, 7, 7, 1, 1}; }
    bool Evaluate(double const *const *p, double *r, double **J) const override;
};
// projectionTwoFrameTwoCamFactor.cpp:43-166   SizedCostFunction<2,7,7,7,7,1,1>
struct ProjTwoFrameTwoCam : CostFunction {
    ProjConst c; double sqrt_info;
    ProjTwoFrameTwoCam(const ProjConst &c_, double si) : c(c_), sqrt_info(si) { num_residuals = 2; block_sizes = {7, 7, 7, 7, 1, 1}; }
    bool Evaluate(double const *const *p, double *r, double **J) const override;
};
// projectionOneFrameTwoCamFactor.cpp:42-134   SizedCostFunction<2,7,7,1,1>
struct ProjOneFrameTwoCam : CostFunction {
    ProjConst c; double sqrt_info;
    ProjOneFrameTwoCam(const ProjConst &c_, double si) : c(c_), sqrt_info(si) { num_residuals = 2; block_sizes = {7, 7, 1, 1}; }
    bool Evaluate(double const *const *p, double *r, double **J) const override;
};

// Public state of IMULegIntegrationBase that the factor reads (imu_leg_integration_base.h:73-85).
struct LegPreintState {
    Mat jacobian{31, 31}, covariance{31, 31};
    double sum_dt = 0;
    V3 delta_p; Quat delta_q; V3 delta_v; V3 delta_epsilon[4];
    V3 linearized_ba, linearized_bg; double linearized_rho[4] = {0, 0, 0, 0};
};

// IMULegIntegrationBase::evaluate, imu_leg_integration_base.cpp:845-898
void imu_leg_residual(const LegPreintState &s, const FactorGlobals &g, V3 Pi, Quat Qi, V3 Vi, V3 Bai, V3 Bgi, const double *rhoi,
                      V3 Pj, Quat Qj, V3 Vj, V3 Baj, V3 Bgj, const double *rhoj, double *residuals31);
// LLT(covariance.inverse()).matrixL().transpose(), imu_leg_factor.cpp:197-198 (31x31 row-major)
bool imu_leg_sqrt_info(const Mat &covariance, Mat &sqrt_info);

// imu_leg_factor.cpp:173-386   SizedCostFunction<31,7,9,4,7,9,4>
struct IMULegFactor : CostFunction {
    const LegPreintState *pre; FactorGlobals g;
    IMULegFactor(const LegPreintState *p, const FactorGlobals &g_) : pre(p), g(g_) { num_residuals = 31; block_sizes = {7, 9, 4, 7, 9, 4}; }
    bool Evaluate(double const *const *p, double *r, double **J) const override;
};

// Public state of IntegrationBase that IMUFactor reads (integration_base.h:200-213); 15-dim error state
// O_P 0, O_R 3, O_V 6, O_BA 9, O_BG 12 (parameters.h:119-126).
struct ImuPreintState {
    Mat jacobian{15, 15}, covariance{15, 15};
    double sum_dt = 0;
    V3 delta_p; Quat delta_q; V3 delta_v;
    V3 linearized_ba, linearized_bg;
};
// IntegrationBase::evaluate, integration_base.h:172-198
void imu_residual(const ImuPreintState &s, const FactorGlobals &g, V3 Pi, Quat Qi, V3 Vi, V3 Bai, V3 Bgi, V3 Pj, Quat Qj, V3 Vj, V3 Baj, V3 Bgj, double *residuals15);
// imu_factor.h:28-188   SizedCostFunction<15,7,9,7,9>
struct IMUFactor : CostFunction {
    const ImuPreintState *pre; FactorGlobals g;
    IMUFactor(const ImuPreintState *p, const FactorGlobals &g_) : pre(p), g(g_) { num_residuals = 15; block_sizes = {7, 9, 7, 9}; }
    bool Evaluate(double const *const *p, double *r, double **J) const override;
};

// State of MarginalizationInfo read by MarginalizationFactor (marginalization_factor.h:76-84)
struct MargInfoLite {
    int n = 0, m = 0;
    std::vector<int> keep_block_size, keep_block_idx;        // idx is absolute (>= m) like the reference
    std::vector<std::vector<double>> keep_block_data;
    Mat linearized_jacobians;                                 // n x n
    std::vector<double> linearized_residuals;                 // n
};
// marginalization_factor.cpp:347-395
struct MarginalizationFactor : CostFunction {
    const MargInfoLite *info;
    explicit MarginalizationFactor(const MargInfoLite *i) : info(i) { num_residuals = i->n; block_sizes = i->keep_block_size; }
    bool Evaluate(double const *const *p, double *r, double **J) const override;
};

// ceres::HuberLoss(a)::Evaluate restated (Ceres 1.14 loss_function.cc): rho[0..2]
inline void huber_loss(double a, double s, double rho[3]) {
    double b = a * a;
    if (s > b) {
        double r = std::sqrt(s);
        rho[0] = 2.0 * a * r - b;
        rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
        rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

// PoseLocalParameterization::Plus, pose_local_parameterization.cpp:12-30
inline void pose_plus(const double *x, const double *delta, double *out) {
    out[0] = x[0] + delta[0]; out[1] = x[1] + delta[1]; out[2] = x[2] + delta[2];
    Quat q(x[6], x[3], x[4], x[5]);
    Quat dq = deltaQ(V3(delta[3], delta[4], delta[5]));
    Quat r = normalized(q * dq);
    out[3] = r.x; out[4] = r.y; out[5] = r.z; out[6] = r.w;
}

}  // namespace oracle
