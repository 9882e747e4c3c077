// oracle/ref_preint.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
#include "ref_preint.h"
#include <cmath>

namespace oracle {

// ---- A1 leg kinematics: closed forms of A1Kinematics.cpp:43-220, re-derived from the FK --------
//   x = ox - lt*s1 - lc*sin(q1+q2)
//   y = oy + d*c0 + lt*c1*s0 + lc*s0*cos(q1+q2)
//   z = d*s0 - lt*c0*c1 - lc*c0*cos(q1+q2)
void a1_fk(const double q[3], double lc, const double f[4], double p[3]) {
    double c0 = std::cos(q[0]), s0 = std::sin(q[0]), c1 = std::cos(q[1]), s1 = std::sin(q[1]);
    double c12 = std::cos(q[1] + q[2]), s12 = std::sin(q[1] + q[2]);
    double ox = f[0], oy = f[1], d = f[2], lt = f[3];
    p[0] = ox - lt * s1 - lc * s12;
    p[1] = oy + d * c0 + lt * c1 * s0 + lc * s0 * c12;
    p[2] = d * s0 - lt * c0 * c1 - lc * c0 * c12;
}
void a1_jac(const double q[3], double lc, const double f[4], double J[9]) {   // column-major J(r,c) = J[c*3+r]
    double c0 = std::cos(q[0]), s0 = std::sin(q[0]), c1 = std::cos(q[1]), s1 = std::sin(q[1]);
    double c12 = std::cos(q[1] + q[2]), s12 = std::sin(q[1] + q[2]);
    double d = f[2], lt = f[3];
    double A = lt * s1 + lc * s12, B = lt * c1 + lc * c12;
    J[0] = 0.0;            J[1] = -d * s0 + c0 * B;  J[2] = d * c0 + s0 * B;
    J[3] = -B;             J[4] = -s0 * A;           J[5] = c0 * A;
    J[6] = -lc * c12;      J[7] = -s0 * lc * s12;    J[8] = c0 * lc * s12;
}
void a1_dfk_drho(const double q[3], double, const double[4], double o[3]) {
    double c12 = std::cos(q[1] + q[2]), s12 = std::sin(q[1] + q[2]);
    o[0] = -s12; o[1] = c12 * std::sin(q[0]); o[2] = -c12 * std::cos(q[0]);
}
void a1_dJ_dq(const double q[3], double lc, const double f[4], double o[27]) {   // 9x3 column-major: o[m*9 + c*3 + r] = dJ(r,c)/dq_m
    double c0 = std::cos(q[0]), s0 = std::sin(q[0]), c1 = std::cos(q[1]), s1 = std::sin(q[1]);
    double c12 = std::cos(q[1] + q[2]), s12 = std::sin(q[1] + q[2]);
    double d = f[2], lt = f[3];
    double A = lt * s1 + lc * s12, B = lt * c1 + lc * c12, Cc = lc * c12, Ss = lc * s12;
    // d/dq0
    o[0] = 0;  o[1] = -d * c0 - s0 * B; o[2] = -d * s0 + c0 * B;
    o[3] = 0;  o[4] = -c0 * A;          o[5] = -s0 * A;
    o[6] = 0;  o[7] = -c0 * Ss;         o[8] = -s0 * Ss;
    // d/dq1
    o[9] = 0;   o[10] = -c0 * A;  o[11] = -s0 * A;
    o[12] = A;  o[13] = -s0 * B;  o[14] = c0 * B;
    o[15] = Ss; o[16] = -s0 * Cc; o[17] = c0 * Cc;
    // d/dq2
    o[18] = 0;  o[19] = -c0 * Ss; o[20] = -s0 * Ss;
    o[21] = Ss; o[22] = -s0 * Cc; o[23] = c0 * Cc;
    o[24] = Ss; o[25] = -s0 * Cc; o[26] = c0 * Cc;
}
void a1_dJ_drho(const double q[3], double, const double[4], double o[9]) {
    double c0 = std::cos(q[0]), s0 = std::sin(q[0]);
    double c12 = std::cos(q[1] + q[2]), s12 = std::sin(q[1] + q[2]);
    o[0] = 0;    o[1] = c0 * c12;  o[2] = s0 * c12;
    o[3] = -c12; o[4] = -s0 * s12; o[5] = c0 * s12;
    o[6] = -c12; o[7] = -s0 * s12; o[8] = c0 * s12;
}

namespace {
enum { ILO_P = 0, ILO_R = 3, ILO_V = 6, ILO_EPS1 = 9, ILO_BA = 21, ILO_BG = 24, ILO_RHO1 = 27 };
enum { ILNO_Ai = 0, ILNO_Gi = 3, ILNO_Ai1 = 6, ILNO_Gi1 = 9, ILNO_BA = 12, ILNO_BG = 15, ILNO_PHIi = 18, ILNO_PHIi1 = 21,
       ILNO_DPHIi = 24, ILNO_DPHIi1 = 27, ILNO_V1 = 30, ILNO_NRHO1 = 42 };
inline M3 colmajor3(const double a[9]) { M3 m; for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) m(r, c) = a[c * 3 + r]; return m; }
// (kron(dphi^T, I3)) * dJ (9 x k, column-major) -> 3 x k : out(r, m) = sum_c dphi[c] * dJ[m*9 + c*3 + r]
inline void kron_mul(const double dphi[3], const double *dJ, int k, double out[3][3]) {
    for (int m = 0; m < k; m++) for (int r = 0; r < 3; r++) {
        double s = 0; for (int c = 0; c < 3; c++) s += dphi[c] * dJ[m * 9 + c * 3 + r]; out[r][m] = s; }
}
}  // namespace

LegPreintegrator::LegPreintegrator(const PreintGlobals &gl, V3 acc_0, V3 gyr_0, const double *phi_0, const double *dphi_0, const double *c_0,
                                   V3 lin_ba, V3 lin_bg, const double *lin_rho) : gl_(gl) {
    acc_0_ = acc_0; gyr_0_ = gyr_0; lin_acc_ = acc_0; lin_gyr_ = gyr_0;
    linearized_ba = lin_ba; linearized_bg = lin_bg;
    sum_dt = 0; delta_p = V3(); delta_q = Quat(); delta_v = V3();
    jacobian = Mat::identity(31); covariance = Mat(31, 31);
    for (int i = 0; i < 12; i++) { phi_0_[i] = lin_phi_[i] = phi_0[i]; dphi_0_[i] = lin_dphi_[i] = dphi_0[i]; }
    for (int i = 0; i < 4; i++) {
        c_0_[i] = lin_c_[i] = c_0[i]; linearized_rho[i] = lin_rho[i];
        foot_force_min_[i] = foot_force_max_[i] = 0; foot_force_contact_threshold_[i] = 0;   // uninitialised in the reference (h:106)
        delta_epsilon[i] = V3(); integration_contact_flag_[i] = true;
        foot_force_window_idx_[i] = 0; foot_force_var_[i] = 0;
        for (int k = 0; k < 5; k++) foot_force_window_[i][k] = 0;
    }
}

void LegPreintegrator::push_back(double dt, V3 acc, V3 gyr, const double *phi, const double *dphi, const double *c) {
    Sample s; s.dt = dt; s.acc = acc; s.gyr = gyr;
    for (int i = 0; i < 12; i++) { s.phi[i] = phi[i]; s.dphi[i] = dphi[i]; }
    for (int i = 0; i < 4; i++) s.c[i] = c[i];
    buf_.push_back(s);
    propagate(dt, acc, gyr, phi, dphi, c);
}

void LegPreintegrator::repropagate(V3 lin_ba, V3 lin_bg, const double *lin_rho) {
    sum_dt = 0; acc_0_ = lin_acc_; gyr_0_ = lin_gyr_;
    for (int i = 0; i < 12; i++) { phi_0_[i] = lin_phi_[i]; dphi_0_[i] = lin_dphi_[i]; }
    for (int i = 0; i < 4; i++) { c_0_[i] = lin_c_[i]; delta_epsilon[i] = V3(); linearized_rho[i] = lin_rho[i]; }
    delta_p = V3(); delta_q = Quat(); delta_v = V3(); sum_delta_epsilon = V3();
    linearized_ba = lin_ba; linearized_bg = lin_bg;
    jacobian = Mat::identity(31); covariance = Mat(31, 31);
    for (auto &s : buf_) propagate(s.dt, s.acc, s.gyr, s.phi, s.dphi, s.c);   // filter state NOT reset (quirk a8')
}

// propagate + midPointIntegration, imu_leg_integration_base.cpp:88-470
void LegPreintegrator::propagate(double _dt, V3 _acc_1, V3 _gyr_1, const double *_phi_1, const double *_dphi_1, const double *_c_1) {
    const V3 _acc_0 = acc_0_, _gyr_0 = gyr_0_;
    const double *_phi_0 = phi_0_, *_dphi_0 = dphi_0_, *_c_0 = c_0_;
    // :152-160
    V3 un_acc_0 = delta_q * (_acc_0 - linearized_ba);
    V3 un_gyr = 0.5 * (_gyr_0 + _gyr_1) - linearized_bg;
    Quat result_delta_q = delta_q * Quat(1, un_gyr.x * _dt / 2, un_gyr.y * _dt / 2, un_gyr.z * _dt / 2);
    V3 un_acc_1 = result_delta_q * (_acc_1 - linearized_ba);
    V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
    V3 result_delta_p = delta_p + delta_v * _dt + 0.5 * un_acc * _dt * _dt;
    V3 result_delta_v = delta_v + un_acc * _dt;

    V3 w_0_x = _gyr_0 - linearized_bg, w_1_x = _gyr_1 - linearized_bg;
    M3 R_w_0_x = skew(w_0_x), R_w_1_x = skew(w_1_x);

    // contact flag, :182-229
    if (gl_.CONTACT_SENSOR_TYPE == 0 || gl_.CONTACT_SENSOR_TYPE == 1) {
        for (int j = 0; j < 4; j++) {
            foot_contact_flag[j] = (_c_1[j] >= 0.5) ? 1 : 0;
            if (foot_contact_flag[j] < 0.5) integration_contact_flag_[j] = false;
        }
    } else if (gl_.CONTACT_SENSOR_TYPE == 2) {
        for (int j = 0; j < 4; j++) {
            double force_mag = 0.5 * (_c_0[j] + _c_1[j]);
            if (force_mag < foot_force_min_[j]) foot_force_min_[j] = 0.9 * foot_force_min_[j] + 0.1 * force_mag;
            if (force_mag > foot_force_max_[j]) foot_force_max_[j] = 0.9 * foot_force_max_[j] + 0.1 * force_mag;
            foot_force_min_[j] *= 0.9991; foot_force_max_[j] *= 0.997;
            foot_force_contact_threshold_[j] = foot_force_min_[j] + gl_.V_N_FORCE_THRES_RATIO * (foot_force_max_[j] - foot_force_min_[j]);
            // assignment into a Vector4i truncates the sigmoid (quirk a8')
            foot_contact_flag[j] = (int)(1.0 / (1 + std::exp(-gl_.V_N_TERM1_STEEP * (force_mag - foot_force_contact_threshold_[j]))));
            foot_force_window_idx_[j]++; foot_force_window_idx_[j] %= 5;
            foot_force_window_[j][foot_force_window_idx_[j]] = force_mag;
            double mean = 0; for (int k = 0; k < 5; k++) mean += foot_force_window_[j][k]; mean /= 5;
            double var = 0; for (int k = 0; k < 5; k++) var += (foot_force_window_[j][k] - mean) * (foot_force_window_[j][k] - mean);
            foot_force_var_[j] = var / 4;
            if (foot_contact_flag[j] < 0.5) integration_contact_flag_[j] = false;
        }
    }

    M3 R0 = toR(delta_q), R1 = toR(result_delta_q);
    const M3 &R_br = gl_.R_br; const V3 p_br = gl_.p_br;
    V3 fi[4], fip1[4], vi[4], vip1[4], result_delta_epsilon[4], gi[4], gip1[4], lo_vel[4];
    M3 Ji[4], Jip1[4], hi[4], hip1[4];
    for (int j = 0; j < 4; j++) {
        double lc = linearized_rho[j];
        double t3[3], t9[9];
        a1_fk(_phi_0 + 3 * j, lc, gl_.rho_fix[j], t3); fi[j] = V3(t3);
        a1_fk(_phi_1 + 3 * j, lc, gl_.rho_fix[j], t3); fip1[j] = V3(t3);
        a1_jac(_phi_0 + 3 * j, lc, gl_.rho_fix[j], t9); Ji[j] = colmajor3(t9);
        a1_jac(_phi_1 + 3 * j, lc, gl_.rho_fix[j], t9); Jip1[j] = colmajor3(t9);
        vi[j] = -(R_br * (Ji[j] * V3(_dphi_0 + 3 * j))) - R_w_0_x * (p_br + R_br * fi[j]);          // :242
        vip1[j] = -(R_br * (Jip1[j] * V3(_dphi_1 + 3 * j))) - R_w_1_x * (p_br + R_br * fip1[j]);  // :243
        result_delta_epsilon[j] = delta_epsilon[j] + 0.5 * (delta_q * vi[j] + result_delta_q * vip1[j]) * _dt;   // :245
        lo_vel[j] = 0.5 * (delta_q * vi[j] + result_delta_q * vip1[j]);
    }
    for (int j = 0; j < 4; j++) {   // :260-286
        double lc = linearized_rho[j];
        double df0[3], df1[3], dJr0[9], dJr1[9], dJq0[27], dJq1[27], k0[3][3], k1[3][3];
        a1_dfk_drho(_phi_0 + 3 * j, lc, gl_.rho_fix[j], df0); a1_dfk_drho(_phi_1 + 3 * j, lc, gl_.rho_fix[j], df1);
        a1_dJ_drho(_phi_0 + 3 * j, lc, gl_.rho_fix[j], dJr0); a1_dJ_drho(_phi_1 + 3 * j, lc, gl_.rho_fix[j], dJr1);
        kron_mul(_dphi_0 + 3 * j, dJr0, 1, k0); kron_mul(_dphi_1 + 3 * j, dJr1, 1, k1);
        gi[j] = -(R0 * (R_br * V3(k0[0][0], k0[1][0], k0[2][0]) + R_w_0_x * (R_br * V3(df0))));
        gip1[j] = -(R1 * (R_br * V3(k1[0][0], k1[1][0], k1[2][0]) + R_w_1_x * (R_br * V3(df1))));
        a1_dJ_dq(_phi_0 + 3 * j, lc, gl_.rho_fix[j], dJq0); a1_dJ_dq(_phi_1 + 3 * j, lc, gl_.rho_fix[j], dJq1);
        kron_mul(_dphi_0 + 3 * j, dJq0, 3, k0); kron_mul(_dphi_1 + 3 * j, dJq1, 3, k1);
        M3 K0, K1m; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { K0(r, c) = k0[r][c]; K1m(r, c) = k1[r][c]; }
        hi[j] = R0 * (R_br * K0 + R_w_0_x * R_br * Ji[j]);
        hip1[j] = R1 * (R_br * K1m + R_w_1_x * R_br * Jip1[j]);
    }
    double uncertainties[12];
    if (gl_.CONTACT_SENSOR_TYPE == 0 || gl_.CONTACT_SENSOR_TYPE == 1) {   // :290-299
        for (int j = 0; j < 4; j++) {
            double n_xy = gl_.V_N_MAX * (1 - foot_contact_flag[j]) + foot_contact_flag[j] * gl_.V_N_MIN_XY;
            double n_z = gl_.V_N_MAX * (1 - foot_contact_flag[j]) + foot_contact_flag[j] * gl_.V_N_MIN_Z;
            uncertainties[3 * j] = n_xy; uncertainties[3 * j + 1] = n_xy; uncertainties[3 * j + 2] = n_z;
        }
    } else {   // :300-317
        for (int j = 0; j < 4; j++) {
            double n1 = gl_.V_N_MAX * (1 - foot_contact_flag[j]) + gl_.V_N_MIN;
            double n2 = gl_.V_N_TERM2_VAR_RESCALE * foot_force_var_[j];
            V3 tmp = lo_vel[j] - delta_v;
            for (int k = 0; k < 3; k++) uncertainties[3 * j + k] = n1 + n2 + gl_.V_N_TERM3_DISTANCE_RESCALE * tmp[k] * tmp[k];
        }
    }
    double rho_uncertainty[4];
    for (int j = 0; j < 4; j++) rho_uncertainty[j] = gl_.RHO_C_N * foot_contact_flag[j] + gl_.RHO_NC_N;   // :319-323
    // weighted average (computed, never consumed by the factor): :325-351
    V3 avg, cnt;
    for (int j = 0; j < 4; j++) for (int k = 0; k < 3; k++) {
        double w = (gl_.V_N_MAX + gl_.V_N_TERM2_VAR_RESCALE + gl_.V_N_TERM3_DISTANCE_RESCALE) / uncertainties[3 * j + k];
        if (w < 0.001) w = 0.001;
        avg[k] += w * lo_vel[j][k] * _dt; cnt[k] += w;
    }
    for (int k = 0; k < 3; k++) avg[k] /= cnt[k];
    sum_delta_epsilon = sum_delta_epsilon + avg;
    if (foot_contact_flag[0] + foot_contact_flag[1] + foot_contact_flag[2] + foot_contact_flag[3] < 1e-6) {   // :354-358
        for (int j = 0; j < 4; j++) rho_uncertainty[j] = gl_.RHO_NC_N;
        for (int k = 0; k < 12; k++) uncertainties[k] = 10e10;
    }
    double N[46];   // :360-374
    {
        double an = gl_.ACC_N * gl_.ACC_N, anz = gl_.ACC_N_Z * gl_.ACC_N_Z, gn = gl_.GYR_N * gl_.GYR_N;
        double aw = gl_.ACC_W * gl_.ACC_W, gw = gl_.GYR_W * gl_.GYR_W, pn = gl_.PHI_N * gl_.PHI_N, dn = gl_.DPHI_N * gl_.DPHI_N;
        double init[30] = {an, an, anz, gn, gn, gn, an, an, anz, gn, gn, gn, aw, aw, aw, gw, gw, gw, pn, pn, pn, pn, pn, pn, dn, dn, dn, dn, dn, dn};
        for (int k = 0; k < 30; k++) N[k] = init[k];
        for (int k = 0; k < 12; k++) N[30 + k] = uncertainties[k];
        for (int k = 0; k < 4; k++) N[42 + k] = rho_uncertainty[k];
    }
    // F, V : :376-465
    V3 w_x = 0.5 * (_gyr_0 + _gyr_1) - linearized_bg;
    V3 a_0_x = _acc_0 - linearized_ba, a_1_x = _acc_1 - linearized_ba;
    M3 R_w_x = skew(w_x), R_a_0_x = skew(a_0_x), R_a_1_x = skew(a_1_x);
    M3 I3 = M3::identity();
    M3 kappa_7 = I3 - R_w_x * _dt;
    Mat F(31, 31), Vm(31, 46);
    F.setBlock(ILO_P, ILO_P, I3);
    M3 kappa_1 = -0.5 * R0 * R_a_0_x * _dt + -0.5 * R1 * R_a_1_x * kappa_7 * _dt;
    F.setBlock(ILO_P, ILO_R, 0.5 * _dt * kappa_1);
    F.setBlock(ILO_P, ILO_V, I3 * _dt);
    F.setBlock(ILO_P, ILO_BA, -0.25 * (R0 + R1) * _dt * _dt);
    F.setBlock(ILO_P, ILO_BG, 0.25 * R1 * R_a_1_x * _dt * _dt * _dt);
    F.setBlock(ILO_R, ILO_R, kappa_7);
    F.setBlock(ILO_R, ILO_BG, -1.0 * I3 * _dt);
    F.setBlock(ILO_V, ILO_R, kappa_1);
    F.setBlock(ILO_V, ILO_V, I3);
    F.setBlock(ILO_V, ILO_BA, -0.5 * (R0 + R1) * _dt);
    F.setBlock(ILO_V, ILO_BG, 0.5 * R1 * R_a_1_x * _dt * _dt);
    for (int j = 0; j < 4; j++) {
        int e = ILO_EPS1 + 3 * j;
        F.setBlock(e, ILO_R, -0.5 * _dt * R0 * skew(vi[j]) - 0.5 * _dt * R1 * skew(vip1[j]) * kappa_7);
        F.setBlock(e, e, I3);
        F.setBlock(e, ILO_BG, 0.5 * _dt * _dt * R1 * skew(vip1[j]) - 0.5 * _dt * (R0 * skew(p_br + R_br * fi[j]) + R1 * skew(p_br + R_br * fip1[j])));
        V3 gcol = 0.5 * _dt * (gi[j] + gip1[j]);
        for (int k = 0; k < 3; k++) F(e + k, ILO_RHO1 + j) = gcol[k];
    }
    F.setBlock(ILO_BA, ILO_BA, I3); F.setBlock(ILO_BG, ILO_BG, I3);
    for (int j = 0; j < 4; j++) F(ILO_RHO1 + j, ILO_RHO1 + j) = 1.0;

    Vm.setBlock(ILO_P, ILNO_Ai, 0.25 * R0 * _dt * _dt);
    M3 VPG = 0.25 * (-R1) * R_a_1_x * _dt * _dt * 0.5 * _dt;
    Vm.setBlock(ILO_P, ILNO_Gi, VPG);
    Vm.setBlock(ILO_P, ILNO_Ai1, 0.25 * R1 * _dt * _dt);
    Vm.setBlock(ILO_P, ILNO_Gi1, VPG);
    Vm.setBlock(ILO_R, ILNO_Gi, 0.5 * I3 * _dt);
    Vm.setBlock(ILO_R, ILNO_Gi1, 0.5 * I3 * _dt);
    Vm.setBlock(ILO_V, ILNO_Ai, 0.5 * R0 * _dt);
    M3 VVG = 0.5 * (-R1) * R_a_1_x * _dt * 0.5 * _dt;
    Vm.setBlock(ILO_V, ILNO_Gi, VVG);
    Vm.setBlock(ILO_V, ILNO_Ai1, 0.5 * R1 * _dt);
    Vm.setBlock(ILO_V, ILNO_Gi1, VVG);
    for (int j = 0; j < 4; j++) {
        int e = ILO_EPS1 + 3 * j;
        Vm.setBlock(e, ILNO_Gi, -0.25 * _dt * _dt * R1 * skew(vip1[j]) + 0.5 * _dt * R0 * skew(p_br + R_br * fi[j]));
        Vm.setBlock(e, ILNO_Gi1, -0.25 * _dt * _dt * R1 * skew(vip1[j]) + 0.5 * _dt * R1 * skew(p_br + R_br * fip1[j]));
        Vm.setBlock(e, ILNO_PHIi, -0.5 * _dt * hi[j]);
        Vm.setBlock(e, ILNO_PHIi1, -0.5 * _dt * hip1[j]);
        Vm.setBlock(e, ILNO_DPHIi, -0.5 * _dt * R0 * R_br * Ji[j]);
        Vm.setBlock(e, ILNO_DPHIi1, -0.5 * _dt * R1 * R_br * Jip1[j]);
        Vm.setBlock(e, ILNO_V1 + 3 * j, -1.0 * I3 * _dt);
    }
    Vm.setBlock(ILO_BA, ILNO_BA, -1.0 * I3 * _dt);
    Vm.setBlock(ILO_BG, ILNO_BG, -1.0 * I3 * _dt);
    for (int j = 0; j < 4; j++) Vm(ILO_RHO1 + j, ILNO_NRHO1 + j) = -_dt;

    jacobian = matmul(F, jacobian);                                       // :467
    Mat FC = matmul(F, covariance);
    Mat cov = matmul(FC, transpose(F));
    Mat VN = Vm;
    for (int i = 0; i < 31; i++) for (int k = 0; k < 46; k++) VN(i, k) *= N[k];
    Mat vnv = matmul(VN, transpose(Vm));
    for (int i = 0; i < 31; i++) for (int k = 0; k < 31; k++) cov(i, k) += vnv(i, k);
    covariance = cov;                                                     // :468

    // propagate() tail, :125-135
    delta_p = result_delta_p; delta_q = normalized(result_delta_q); delta_v = result_delta_v;
    for (int j = 0; j < 4; j++) delta_epsilon[j] = result_delta_epsilon[j];
    sum_dt += _dt;
    acc_0_ = _acc_1; gyr_0_ = _gyr_1;
    for (int i = 0; i < 12; i++) { phi_0_[i] = _phi_1[i]; dphi_0_[i] = _dphi_1[i]; }
    for (int i = 0; i < 4; i++) c_0_[i] = _c_1[i];
}

// ---- IntegrationBase::{push_back, propagate, midPointIntegration}, integration_base.h:40-170 --------------------
ImuPreintegrator::ImuPreintegrator(const PreintGlobals &gl, V3 acc_0, V3 gyr_0, V3 lin_ba, V3 lin_bg) : gl_(gl), acc_0_(acc_0), gyr_0_(gyr_0) {
    linearized_ba = lin_ba; linearized_bg = lin_bg; sum_dt = 0; delta_q = Quat();
    jacobian = Mat::identity(15); covariance = Mat(15, 15);
}
void ImuPreintegrator::push_back(double _dt, V3 _acc_1, V3 _gyr_1) {
    const V3 _acc_0 = acc_0_, _gyr_0 = gyr_0_;
    V3 un_acc_0 = delta_q * (_acc_0 - linearized_ba);
    V3 un_gyr = 0.5 * (_gyr_0 + _gyr_1) - linearized_bg;
    Quat result_delta_q = delta_q * Quat(1, un_gyr.x * _dt / 2, un_gyr.y * _dt / 2, un_gyr.z * _dt / 2);
    V3 un_acc_1 = result_delta_q * (_acc_1 - linearized_ba);
    V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
    V3 result_delta_p = delta_p + delta_v * _dt + 0.5 * un_acc * _dt * _dt;
    V3 result_delta_v = delta_v + un_acc * _dt;
    M3 R_w_x = skew(un_gyr), R_a_0_x = skew(_acc_0 - linearized_ba), R_a_1_x = skew(_acc_1 - linearized_ba);
    M3 R0 = toR(delta_q), R1 = toR(result_delta_q), I3 = M3::identity();
    Mat F(15, 15), V(15, 18);
    F.setBlock(0, 0, I3);
    F.setBlock(0, 3, -0.25 * R0 * R_a_0_x * _dt * _dt + -0.25 * R1 * R_a_1_x * (I3 - R_w_x * _dt) * _dt * _dt);
    F.setBlock(0, 6, I3 * _dt);
    F.setBlock(0, 9, -0.25 * (R0 + R1) * _dt * _dt);
    F.setBlock(0, 12, -0.25 * R1 * R_a_1_x * _dt * _dt * -_dt);
    F.setBlock(3, 3, I3 - R_w_x * _dt);
    F.setBlock(3, 12, -1.0 * I3 * _dt);
    F.setBlock(6, 3, -0.5 * R0 * R_a_0_x * _dt + -0.5 * R1 * R_a_1_x * (I3 - R_w_x * _dt) * _dt);
    F.setBlock(6, 6, I3);
    F.setBlock(6, 9, -0.5 * (R0 + R1) * _dt);
    F.setBlock(6, 12, -0.5 * R1 * R_a_1_x * _dt * -_dt);
    F.setBlock(9, 9, I3); F.setBlock(12, 12, I3);
    V.setBlock(0, 0, 0.25 * R0 * _dt * _dt);
    M3 v03 = 0.25 * (-R1) * R_a_1_x * _dt * _dt * 0.5 * _dt;
    V.setBlock(0, 3, v03); V.setBlock(0, 6, 0.25 * R1 * _dt * _dt); V.setBlock(0, 9, v03);
    V.setBlock(3, 3, 0.5 * I3 * _dt); V.setBlock(3, 9, 0.5 * I3 * _dt);
    V.setBlock(6, 0, 0.5 * R0 * _dt);
    M3 v63 = 0.5 * (-R1) * R_a_1_x * _dt * 0.5 * _dt;
    V.setBlock(6, 3, v63); V.setBlock(6, 6, 0.5 * R1 * _dt); V.setBlock(6, 9, v63);
    V.setBlock(9, 12, I3 * _dt); V.setBlock(12, 15, I3 * _dt);
    double N[18];
    for (int k = 0; k < 3; k++) { N[k] = N[6 + k] = gl_.ACC_N * gl_.ACC_N; N[3 + k] = N[9 + k] = gl_.GYR_N * gl_.GYR_N; N[12 + k] = gl_.ACC_W * gl_.ACC_W; N[15 + k] = gl_.GYR_W * gl_.GYR_W; }
    jacobian = matmul(F, jacobian);
    Mat cov = matmul(matmul(F, covariance), transpose(F));
    Mat VN = V; for (int i = 0; i < 15; i++) for (int k = 0; k < 18; k++) VN(i, k) *= N[k];
    Mat vnv = matmul(VN, transpose(V));
    for (int i = 0; i < 15; i++) for (int k = 0; k < 15; k++) cov(i, k) += vnv(i, k);
    covariance = cov;
    delta_p = result_delta_p; delta_q = normalized(result_delta_q); delta_v = result_delta_v;
    sum_dt += _dt; acc_0_ = _acc_1; gyr_0_ = _gyr_1;
}

}  // namespace oracle
