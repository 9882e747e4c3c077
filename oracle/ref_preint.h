// oracle/ref_preint.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// CPU restatement of IMULegIntegrationBase (src/factor/imu_leg_integration_base.{h,cpp}) and
// A1Kinematics (src/legKinematics/A1Kinematics.cpp).  PARITY UNPINNED.
#pragma once
#include "ref_factors.h"

namespace oracle {

// A1Kinematics.cpp:43-220 -- closed forms; in1=q(3), lc = rho_opt, in3 = rho_fix = [ox, oy, d, lt].
// Outputs column-major like the Eigen matrices the reference fills through .data().
void a1_fk(const double q[3], double lc, const double rho_fix[4], double p_bf[3]);
void a1_jac(const double q[3], double lc, const double rho_fix[4], double jac[9]);
void a1_dfk_drho(const double q[3], double lc, const double rho_fix[4], double out[3]);
void a1_dJ_dq(const double q[3], double lc, const double rho_fix[4], double out[27]);
void a1_dJ_drho(const double q[3], double lc, const double rho_fix[4], double out[9]);

// Globals read by the preintegrator (parameters.h:59-75).
struct PreintGlobals {
    double ACC_N = 0.9, ACC_N_Z = 2.5, GYR_N = 0.05, ACC_W = 0.0004, GYR_W = 0.0002;
    double PHI_N = 1e-5, DPHI_N = 1e-5, RHO_C_N = 1e-8, RHO_NC_N = 1e-11;
    double V_N_MIN_XY = 1e-3, V_N_MIN_Z = 5e-3, V_N_MIN = 5e-3, V_N_MAX = 900.0;
    double V_N_FORCE_THRES_RATIO = 0.8, V_N_TERM1_STEEP = 10, V_N_TERM2_VAR_RESCALE = 1e-6, V_N_TERM3_DISTANCE_RESCALE = 1e-3;
    int CONTACT_SENSOR_TYPE = 0;
    double rho_fix[4][4];
    V3 p_br; M3 R_br;
};

class LegPreintegrator : public LegPreintState {
public:
    // imu_leg_integration_base.cpp:7-47
    LegPreintegrator(const PreintGlobals &gl, V3 acc_0, V3 gyr_0, const double *phi_0, const double *dphi_0, const double *c_0,
                     V3 lin_ba, V3 lin_bg, const double *lin_rho);
    // :49-59
    void push_back(double dt, V3 acc, V3 gyr, const double *phi, const double *dphi, const double *c);
    // :62-86
    void repropagate(V3 lin_ba, V3 lin_bg, const double *lin_rho);

    int foot_contact_flag[4] = {0, 0, 0, 0};   // Vector4i in the reference (h:84): sigmoid truncates to 0/1
    V3 sum_delta_epsilon;
private:
    void propagate(double dt, V3 acc_1, V3 gyr_1, const double *phi_1, const double *dphi_1, const double *c_1);  // :88-136
    PreintGlobals gl_;
    V3 acc_0_, gyr_0_; double phi_0_[12], dphi_0_[12], c_0_[4];
    V3 lin_acc_, lin_gyr_; double lin_phi_[12], lin_dphi_[12], lin_c_[4];
    double foot_force_min_[4], foot_force_max_[4], foot_force_contact_threshold_[4];
    double foot_force_window_[4][5]; int foot_force_window_idx_[4]; double foot_force_var_[4];
    bool integration_contact_flag_[4];
    struct Sample { double dt; V3 acc, gyr; double phi[12], dphi[12], c[4]; };
    std::vector<Sample> buf_;
};

// IntegrationBase (integration_base.h:22-170): plain IMU midpoint preintegration, 15 x 15 F, 15 x 18 V
class ImuPreintegrator : public ImuPreintState {
public:
    ImuPreintegrator(const PreintGlobals &gl, V3 acc_0, V3 gyr_0, V3 lin_ba, V3 lin_bg);
    void push_back(double dt, V3 acc, V3 gyr);
private:
    PreintGlobals gl_; V3 acc_0_, gyr_0_;
};

}  // namespace oracle
