// oracle/ref_factors.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Statement-by-statement restatement of the reference factor Evaluate() bodies.
#include "ref_factors.h"
#include <limits>
#include <algorithm>

namespace oracle {

namespace {
struct R23 { double m[2][3]; };
// reduce (2x3) * jaco(3x6 given as two 3x3 halves) -> row-major 2x7 with zero last column
inline void write_pose_jac(double *J, const R23 &reduce, const M3 &left, const M3 &right) {
    for (int r = 0; r < 2; r++) {
        for (int c = 0; c < 3; c++) {
            double sl = 0, sr = 0;
            for (int k = 0; k < 3; k++) { sl += reduce.m[r][k] * left(k, c); sr += reduce.m[r][k] * right(k, c); }
            J[r * 7 + c] = sl; J[r * 7 + 3 + c] = sr;
        }
        J[r * 7 + 6] = 0.0;
    }
}
inline void reduce_times_vec(const R23 &reduce, V3 v, double out[2]) {
    for (int r = 0; r < 2; r++) out[r] = reduce.m[r][0] * v.x + reduce.m[r][1] * v.y + reduce.m[r][2] * v.z;
}
inline R23 make_reduce(V3 pc, double sqrt_info) {
    double dep = pc.z;
    R23 red;
    red.m[0][0] = 1. / dep; red.m[0][1] = 0; red.m[0][2] = -pc.x / (dep * dep);
    red.m[1][0] = 0; red.m[1][1] = 1. / dep; red.m[1][2] = -pc.y / (dep * dep);
    for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) red.m[r][c] *= sqrt_info;  // sqrt_info = s * I2
    return red;
}
}  // namespace

// ---- projectionTwoFrameOneCamFactor.cpp:43-150 -------------------------------------------------
bool ProjTwoFrameOneCam::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
    V3 Pi(p[0]); Quat Qi(p[0][6], p[0][3], p[0][4], p[0][5]);
    V3 Pj(p[1]); Quat Qj(p[1][6], p[1][3], p[1][4], p[1][5]);
    V3 tic(p[2]); Quat qic(p[2][6], p[2][3], p[2][4], p[2][5]);
    double inv_dep_i = p[3][0];
    double td = p[4][0];

    V3 pts_i_td = c.pts_i - (td - c.td_i) * c.velocity_i;          // :60
    V3 pts_j_td = c.pts_j - (td - c.td_j) * c.velocity_j;          // :61
    V3 pts_camera_i = pts_i_td / inv_dep_i;                        // :62
    V3 pts_imu_i = qic * pts_camera_i + tic;
    V3 pts_w = Qi * pts_imu_i + Pi;
    V3 pts_imu_j = inverse(Qj) * (pts_w - Pj);
    V3 pts_camera_j = inverse(qic) * (pts_imu_j - tic);            // :66
    double dep_j = pts_camera_j.z;
    residuals[0] = sqrt_info * (pts_camera_j.x / dep_j - pts_j_td.x);   // :72-76
    residuals[1] = sqrt_info * (pts_camera_j.y / dep_j - pts_j_td.y);

    if (jacobians) {
        M3 Ri = toR(Qi), Rj = toR(Qj), ric = toR(qic);
        R23 reduce = make_reduce(pts_camera_j, sqrt_info);
        M3 ricT = transpose(ric), RjT = transpose(Rj);
        if (jacobians[0]) {   // :101-111
            M3 left = ricT * RjT;
            M3 right = ricT * RjT * Ri * (-skew(pts_imu_i));
            write_pose_jac(jacobians[0], reduce, left, right);
        }
        if (jacobians[1]) {   // :113-123
            M3 left = ricT * (-RjT);
            M3 right = ricT * skew(pts_imu_j);
            write_pose_jac(jacobians[1], reduce, left, right);
        }
        if (jacobians[2]) {   // :124-134
            M3 left = ricT * (RjT * Ri - M3::identity());
            M3 tmp_r = ricT * RjT * Ri * ric;
            M3 right = -(tmp_r * skew(pts_camera_i)) + skew(tmp_r * pts_camera_i) +
                       skew(ricT * (RjT * (Ri * tic + Pi - Pj) - tic));
            write_pose_jac(jacobians[2], reduce, left, right);
        }
        if (jacobians[3]) {   // :135-139
            V3 v = (ricT * RjT * Ri * ric) * pts_i_td * -1.0 / (inv_dep_i * inv_dep_i);
            reduce_times_vec(reduce, v, jacobians[3]);
        }
        if (jacobians[4]) {   // :140-145
            V3 v = (ricT * RjT * Ri * ric) * c.velocity_i / inv_dep_i * -1.0;
            double t[2]; reduce_times_vec(reduce, v, t);
            jacobians[4][0] = t[0] + sqrt_info * c.velocity_j.x;
            jacobians[4][1] = t[1] + sqrt_info * c.velocity_j.y;
        }
    }
    return true;
}

// ---- projectionTwoFrameTwoCamFactor.cpp:43-166 -------------------------------------------------
bool ProjTwoFrameTwoCam::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
    V3 Pi(p[0]); Quat Qi(p[0][6], p[0][3], p[0][4], p[0][5]);
    V3 Pj(p[1]); Quat Qj(p[1][6], p[1][3], p[1][4], p[1][5]);
    V3 tic(p[2]); Quat qic(p[2][6], p[2][3], p[2][4], p[2][5]);
    V3 tic2(p[3]); Quat qic2(p[3][6], p[3][3], p[3][4], p[3][5]);
    double inv_dep_i = p[4][0];
    double td = p[5][0];

    V3 pts_i_td = c.pts_i - (td - c.td_i) * c.velocity_i;
    V3 pts_j_td = c.pts_j - (td - c.td_j) * c.velocity_j;
    V3 pts_camera_i = pts_i_td / inv_dep_i;
    V3 pts_imu_i = qic * pts_camera_i + tic;
    V3 pts_w = Qi * pts_imu_i + Pi;
    V3 pts_imu_j = inverse(Qj) * (pts_w - Pj);
    V3 pts_camera_j = inverse(qic2) * (pts_imu_j - tic2);          // :70
    double dep_j = pts_camera_j.z;
    residuals[0] = sqrt_info * (pts_camera_j.x / dep_j - pts_j_td.x);
    residuals[1] = sqrt_info * (pts_camera_j.y / dep_j - pts_j_td.y);

    if (jacobians) {
        M3 Ri = toR(Qi), Rj = toR(Qj), ric = toR(qic), ric2 = toR(qic2);
        R23 reduce = make_reduce(pts_camera_j, sqrt_info);
        M3 ric2T = transpose(ric2), RjT = transpose(Rj);
        if (jacobians[0]) {   // :107-117
            M3 left = ric2T * RjT;
            M3 right = ric2T * RjT * Ri * (-skew(pts_imu_i));
            write_pose_jac(jacobians[0], reduce, left, right);
        }
        if (jacobians[1]) {   // :119-128
            M3 left = ric2T * (-RjT);
            M3 right = ric2T * skew(pts_imu_j);
            write_pose_jac(jacobians[1], reduce, left, right);
        }
        if (jacobians[2]) {   // :129-137
            M3 left = ric2T * RjT * Ri;
            M3 right = ric2T * RjT * Ri * ric * (-skew(pts_camera_i));
            write_pose_jac(jacobians[2], reduce, left, right);
        }
        if (jacobians[3]) {   // :138-146
            M3 left = -ric2T;
            M3 right = skew(pts_camera_j);
            write_pose_jac(jacobians[3], reduce, left, right);
        }
        if (jacobians[4]) {   // :147-155
            V3 v = (ric2T * RjT * Ri * ric) * pts_i_td * -1.0 / (inv_dep_i * inv_dep_i);
            reduce_times_vec(reduce, v, jacobians[4]);
        }
        if (jacobians[5]) {   // :156-161
            V3 v = (ric2T * RjT * Ri * ric) * c.velocity_i / inv_dep_i * -1.0;
            double t[2]; reduce_times_vec(reduce, v, t);
            jacobians[5][0] = t[0] + sqrt_info * c.velocity_j.x;
            jacobians[5][1] = t[1] + sqrt_info * c.velocity_j.y;
        }
    }
    return true;
}

// ---- projectionOneFrameTwoCamFactor.cpp:42-134 -------------------------------------------------
bool ProjOneFrameTwoCam::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
    V3 tic(p[0]); Quat qic(p[0][6], p[0][3], p[0][4], p[0][5]);
    V3 tic2(p[1]); Quat qic2(p[1][6], p[1][3], p[1][4], p[1][5]);
    double inv_dep_i = p[2][0];
    double td = p[3][0];

    V3 pts_i_td = c.pts_i - (td - c.td_i) * c.velocity_i;
    V3 pts_j_td = c.pts_j - (td - c.td_j) * c.velocity_j;
    V3 pts_camera_i = pts_i_td / inv_dep_i;
    V3 pts_imu_i = qic * pts_camera_i + tic;
    V3 pts_imu_j = pts_imu_i;
    V3 pts_camera_j = inverse(qic2) * (pts_imu_j - tic2);
    double dep_j = pts_camera_j.z;
    residuals[0] = sqrt_info * (pts_camera_j.x / dep_j - pts_j_td.x);
    residuals[1] = sqrt_info * (pts_camera_j.y / dep_j - pts_j_td.y);

    if (jacobians) {
        M3 ric = toR(qic), ric2 = toR(qic2);
        R23 reduce = make_reduce(pts_camera_j, sqrt_info);
        M3 ric2T = transpose(ric2);
        if (jacobians[0]) {   // :98-106
            M3 left = ric2T;
            M3 right = ric2T * ric * (-skew(pts_camera_i));
            write_pose_jac(jacobians[0], reduce, left, right);
        }
        if (jacobians[1]) {   // :107-115
            M3 left = -ric2T;
            M3 right = skew(pts_camera_j);
            write_pose_jac(jacobians[1], reduce, left, right);
        }
        if (jacobians[2]) {   // :116-124  NB: pts_i, not pts_i_td (quirk kept)
            V3 v = (ric2T * ric) * c.pts_i * -1.0 / (inv_dep_i * inv_dep_i);
            reduce_times_vec(reduce, v, jacobians[2]);
        }
        if (jacobians[3]) {   // :125-130
            V3 v = (ric2T * ric) * c.velocity_i / inv_dep_i * -1.0;
            double t[2]; reduce_times_vec(reduce, v, t);
            jacobians[3][0] = t[0] + sqrt_info * c.velocity_j.x;
            jacobians[3][1] = t[1] + sqrt_info * c.velocity_j.y;
        }
    }
    return true;
}

// ---- IMULegIntegrationBase::evaluate, imu_leg_integration_base.cpp:845-898 ---------------------
// ILStateOrder (parameters.h:135-150): P0 R3 V6 EPS1..4 @9,12,15,18 BA21 BG24 RHO1..4 @27..30
enum { ILO_P = 0, ILO_R = 3, ILO_V = 6, ILO_EPS1 = 9, ILO_BA = 21, ILO_BG = 24, ILO_RHO1 = 27 };

void imu_leg_residual(const LegPreintState &s, const FactorGlobals &g, V3 Pi, Quat Qi, V3 Vi, V3 Bai, V3 Bgi, const double *rhoi,
                      V3 Pj, Quat Qj, V3 Vj, V3 Baj, V3 Bgj, const double *rhoj, double *res) {
    M3 dp_dba = s.jacobian.block3(ILO_P, ILO_BA), dp_dbg = s.jacobian.block3(ILO_P, ILO_BG);
    M3 dq_dbg = s.jacobian.block3(ILO_R, ILO_BG);
    M3 dv_dba = s.jacobian.block3(ILO_V, ILO_BA), dv_dbg = s.jacobian.block3(ILO_V, ILO_BG);
    V3 dba = Bai - s.linearized_ba, dbg = Bgi - s.linearized_bg;
    Quat corrected_delta_q = s.delta_q * deltaQ(dq_dbg * dbg);
    V3 corrected_delta_v = s.delta_v + dv_dba * dba + dv_dbg * dbg;
    V3 corrected_delta_p = s.delta_p + dp_dba * dba + dp_dbg * dbg;
    double sum_dt = s.sum_dt;
    V3 rp = inverse(Qi) * (0.5 * g.G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
    V3 rq = 2.0 * (inverse(corrected_delta_q) * (inverse(Qi) * Qj)).vec();
    V3 rv = inverse(Qi) * (g.G * sum_dt + Vj - Vi) - corrected_delta_v;
    for (int k = 0; k < 3; k++) { res[ILO_P + k] = rp[k]; res[ILO_R + k] = rq[k]; res[ILO_V + k] = rv[k]; }
    for (int j = 0; j < 4; j++) {
        M3 dep_dbg = s.jacobian.block3(ILO_EPS1 + 3 * j, ILO_BG);
        V3 dep_drho(s.jacobian(ILO_EPS1 + 3 * j, ILO_RHO1 + j), s.jacobian(ILO_EPS1 + 3 * j + 1, ILO_RHO1 + j), s.jacobian(ILO_EPS1 + 3 * j + 2, ILO_RHO1 + j));
        double drho = rhoi[j] - s.linearized_rho[j];
        V3 corrected = s.delta_epsilon[j] + dep_dbg * dbg + dep_drho * drho;
        V3 re = inverse(Qi) * (Pj - Pi) - corrected;
        for (int k = 0; k < 3; k++) res[ILO_EPS1 + 3 * j + k] = re[k];
        res[ILO_RHO1 + j] = rhoj[j] - rhoi[j];
    }
    V3 rba = Baj - Bai, rbg = Bgj - Bgi;
    for (int k = 0; k < 3; k++) { res[ILO_BA + k] = rba[k]; res[ILO_BG + k] = rbg[k]; }
}

bool imu_leg_sqrt_info(const Mat &covariance, Mat &sqrt_info) {
    Mat inv;
    if (!inverse_partial_piv_lu(covariance, inv)) return false;
    Mat L = inv;                      // LLT reads the lower triangle only
    if (!cholesky_lower(L)) return false;
    int n = covariance.r;
    sqrt_info = Mat(n, n);
    for (int i = 0; i < n; i++) for (int j = i; j < n; j++) sqrt_info(i, j) = L(j, i);   // matrixL().transpose()
    return true;
}

// ---- IMULegFactor::Evaluate, imu_leg_factor.cpp:173-386 ----------------------------------------
bool IMULegFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
    const LegPreintState &s = *pre;
    V3 Pi(p[0]); Quat Qi(p[0][6], p[0][3], p[0][4], p[0][5]);
    V3 Vi(p[1]), Bai(p[1] + 3), Bgi(p[1] + 6);
    const double *rhoi = p[2];
    V3 Pj(p[3]); Quat Qj(p[3][6], p[3][3], p[3][4], p[3][5]);
    V3 Vj(p[4]), Baj(p[4] + 3), Bgj(p[4] + 6);
    const double *rhoj = p[5];

    double raw[31];
    imu_leg_residual(s, g, Pi, Qi, Vi, Bai, Bgi, rhoi, Pj, Qj, Vj, Baj, Bgj, rhoj, raw);
    Mat sqrt_info;
    if (!imu_leg_sqrt_info(s.covariance, sqrt_info)) return false;       // :197-198, every call
    for (int i = 0; i < 31; i++) { double t = 0; for (int k = i; k < 31; k++) t += sqrt_info(i, k) * raw[k]; residuals[i] = t; }

    if (jacobians) {
        double sum_dt = s.sum_dt;
        M3 dp_dba = s.jacobian.block3(ILO_P, ILO_BA), dp_dbg = s.jacobian.block3(ILO_P, ILO_BG);
        M3 dq_dbg = s.jacobian.block3(ILO_R, ILO_BG);
        M3 dv_dba = s.jacobian.block3(ILO_V, ILO_BA), dv_dbg = s.jacobian.block3(ILO_V, ILO_BG);
        M3 RiT = toR(inverse(Qi));                                        // Qi.inverse().toRotationMatrix()
        auto whiten_and_store = [&](const Mat &Jm, double *out) {         // out = sqrt_info * Jm, row-major
            int cols = Jm.c;
            for (int i = 0; i < 31; i++) for (int cc = 0; cc < cols; cc++) {
                double t = 0; for (int k = i; k < 31; k++) t += sqrt_info(i, k) * Jm(k, cc);
                out[i * cols + cc] = t;
            }
        };
        if (jacobians[0]) {   // :221-252
            Mat J(31, 7);
            J.setBlock(ILO_P, 0, -RiT);
            J.setBlock(ILO_P, 3, skew(inverse(Qi) * (0.5 * g.G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
            Quat corrected_delta_q = s.delta_q * deltaQ(dq_dbg * (Bgi - s.linearized_bg));
            J.setBlock(ILO_R, 3, -QleftQrightBR(inverse(Qj) * Qi, corrected_delta_q));
            J.setBlock(ILO_V, 3, skew(inverse(Qi) * (g.G * sum_dt + Vj - Vi)));
            for (int j = 0; j < 4; j++) {
                J.setBlock(ILO_EPS1 + 3 * j, 0, -RiT);
                J.setBlock(ILO_EPS1 + 3 * j, 3, skew(inverse(Qi) * (Pj - Pi)));
            }
            whiten_and_store(J, jacobians[0]);
        }
        if (jacobians[1]) {   // :254-295
            Mat J(31, 9);
            J.setBlock(ILO_P, 0, -1.0 * RiT * sum_dt);
            J.setBlock(ILO_P, 3, -dp_dba);
            J.setBlock(ILO_P, 6, -dp_dbg);
            J.setBlock(ILO_R, 6, -(QleftBR(inverse(Qj) * Qi * s.delta_q) * dq_dbg));
            J.setBlock(ILO_V, 0, -RiT);
            J.setBlock(ILO_V, 3, -dv_dba);
            J.setBlock(ILO_V, 6, -dv_dbg);
            for (int j = 0; j < 4; j++) J.setBlock(ILO_EPS1 + 3 * j, 6, -s.jacobian.block3(ILO_EPS1 + 3 * j, ILO_BG));
            J.setBlock(ILO_BA, 3, -M3::identity());
            J.setBlock(ILO_BG, 6, -M3::identity());
            whiten_and_store(J, jacobians[1]);
        }
        if (jacobians[2]) {   // :297-318
            Mat J(31, 4);
            for (int j = 0; j < 4; j++) {
                for (int k = 0; k < 3; k++) J(ILO_EPS1 + 3 * j + k, j) = -s.jacobian(ILO_EPS1 + 3 * j + k, ILO_RHO1 + j);
                J(ILO_RHO1 + j, j) = -1.0;
            }
            whiten_and_store(J, jacobians[2]);
        }
        if (jacobians[3]) {   // :320-344
            Mat J(31, 7);
            J.setBlock(ILO_P, 0, RiT);
            Quat corrected_delta_q = s.delta_q * deltaQ(dq_dbg * (Bgi - s.linearized_bg));
            J.setBlock(ILO_R, 3, QleftBR(inverse(corrected_delta_q) * inverse(Qi) * Qj));
            for (int j = 0; j < 4; j++) J.setBlock(ILO_EPS1 + 3 * j, 0, RiT);
            whiten_and_store(J, jacobians[3]);
        }
        if (jacobians[4]) {   // :346-365
            Mat J(31, 9);
            J.setBlock(ILO_V, 0, RiT);
            J.setBlock(ILO_BA, 3, M3::identity());
            J.setBlock(ILO_BG, 6, M3::identity());
            whiten_and_store(J, jacobians[4]);
        }
        if (jacobians[5]) {   // :366-383
            Mat J(31, 4);
            for (int j = 0; j < 4; j++) J(ILO_RHO1 + j, j) = 1.0;
            whiten_and_store(J, jacobians[5]);
        }
    }
    return true;
}

// ---- IntegrationBase::evaluate (integration_base.h:172-198) and IMUFactor::Evaluate (imu_factor.h:28-188) -------
enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };
void imu_residual(const ImuPreintState &s, const FactorGlobals &g, V3 Pi, Quat Qi, V3 Vi, V3 Bai, V3 Bgi, V3 Pj, Quat Qj, V3 Vj, V3 Baj, V3 Bgj, double *res) {
    M3 dp_dba = s.jacobian.block3(O_P, O_BA), dp_dbg = s.jacobian.block3(O_P, O_BG), dq_dbg = s.jacobian.block3(O_R, O_BG);
    M3 dv_dba = s.jacobian.block3(O_V, O_BA), dv_dbg = s.jacobian.block3(O_V, O_BG);
    V3 dba = Bai - s.linearized_ba, dbg = Bgi - s.linearized_bg;
    Quat corrected_delta_q = s.delta_q * deltaQ(dq_dbg * dbg);
    V3 corrected_delta_v = s.delta_v + dv_dba * dba + dv_dbg * dbg;
    V3 corrected_delta_p = s.delta_p + dp_dba * dba + dp_dbg * dbg;
    const double sum_dt = s.sum_dt;
    V3 rp = inverse(Qi) * (0.5 * g.G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
    V3 rq = 2.0 * (inverse(corrected_delta_q) * (inverse(Qi) * Qj)).vec();
    V3 rv = inverse(Qi) * (g.G * sum_dt + Vj - Vi) - corrected_delta_v;
    V3 rba = Baj - Bai, rbg = Bgj - Bgi;
    for (int k = 0; k < 3; k++) { res[O_P + k] = rp[k]; res[O_R + k] = rq[k]; res[O_V + k] = rv[k]; res[O_BA + k] = rba[k]; res[O_BG + k] = rbg[k]; }
}

bool IMUFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
    const ImuPreintState &s = *pre;
    V3 Pi(p[0]); Quat Qi(p[0][6], p[0][3], p[0][4], p[0][5]);
    V3 Vi(p[1]), Bai(p[1] + 3), Bgi(p[1] + 6);
    V3 Pj(p[2]); Quat Qj(p[2][6], p[2][3], p[2][4], p[2][5]);
    V3 Vj(p[3]), Baj(p[3] + 3), Bgj(p[3] + 6);
    double raw[15];
    imu_residual(s, g, Pi, Qi, Vi, Bai, Bgi, Pj, Qj, Vj, Baj, Bgj, raw);
    Mat sqrt_info;
    if (!imu_leg_sqrt_info(s.covariance, sqrt_info)) return false;          // imu_factor.h:73, every call
    for (int i = 0; i < 15; i++) { double t = 0; for (int k = i; k < 15; k++) t += sqrt_info(i, k) * raw[k]; residuals[i] = t; }
    if (jacobians) {
        const double sum_dt = s.sum_dt;
        M3 dp_dba = s.jacobian.block3(O_P, O_BA), dp_dbg = s.jacobian.block3(O_P, O_BG), dq_dbg = s.jacobian.block3(O_R, O_BG);
        M3 dv_dba = s.jacobian.block3(O_V, O_BA), dv_dbg = s.jacobian.block3(O_V, O_BG);
        M3 RiT = toR(inverse(Qi));
        auto whiten_and_store = [&](const Mat &Jm, double *out) {
            for (int i = 0; i < 15; i++) for (int cc = 0; cc < Jm.c; cc++) { double t = 0; for (int k = i; k < 15; k++) t += sqrt_info(i, k) * Jm(k, cc); out[i * Jm.c + cc] = t; }
        };
        Quat corrected_delta_q = s.delta_q * deltaQ(dq_dbg * (Bgi - s.linearized_bg));
        if (jacobians[0]) {   // imu_factor.h:95-123
            Mat J(15, 7);
            J.setBlock(O_P, O_P, -RiT);
            J.setBlock(O_P, O_R, skew(inverse(Qi) * (0.5 * g.G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
            J.setBlock(O_R, O_R, -QleftQrightBR(inverse(Qj) * Qi, corrected_delta_q));
            J.setBlock(O_V, O_R, skew(inverse(Qi) * (g.G * sum_dt + Vj - Vi)));
            whiten_and_store(J, jacobians[0]);
        }
        if (jacobians[1]) {   // :124-152
            Mat J(15, 9);
            J.setBlock(O_P, 0, -1.0 * RiT * sum_dt); J.setBlock(O_P, 3, -dp_dba); J.setBlock(O_P, 6, -dp_dbg);
            J.setBlock(O_R, 6, -(QleftBR(inverse(Qj) * Qi * s.delta_q) * dq_dbg));
            J.setBlock(O_V, 0, -RiT); J.setBlock(O_V, 3, -dv_dba); J.setBlock(O_V, 6, -dv_dbg);
            J.setBlock(O_BA, 3, -M3::identity()); J.setBlock(O_BG, 6, -M3::identity());
            whiten_and_store(J, jacobians[1]);
        }
        if (jacobians[2]) {   // :153-171
            Mat J(15, 7);
            J.setBlock(O_P, O_P, RiT);
            J.setBlock(O_R, O_R, QleftBR(inverse(corrected_delta_q) * inverse(Qi) * Qj));
            whiten_and_store(J, jacobians[2]);
        }
        if (jacobians[3]) {   // :172-184
            Mat J(15, 9);
            J.setBlock(O_V, 0, RiT); J.setBlock(O_BA, 3, M3::identity()); J.setBlock(O_BG, 6, M3::identity());
            whiten_and_store(J, jacobians[3]);
        }
    }
    return true;
}

// ---- MarginalizationFactor::Evaluate, marginalization_factor.cpp:347-395 -----------------------
bool MarginalizationFactor::Evaluate(double const *const *p, double *residuals, double **jacobians) const {
    int n = info->n, m = info->m;
    std::vector<double> dx(n, 0.0);
    for (size_t i = 0; i < info->keep_block_size.size(); i++) {
        int size = info->keep_block_size[i];
        int idx = info->keep_block_idx[i] - m;
        const double *x = p[i];
        const double *x0 = info->keep_block_data[i].data();
        if (size != 7) {
            for (int k = 0; k < size; k++) dx[idx + k] = x[k] - x0[k];
        } else {
            for (int k = 0; k < 3; k++) dx[idx + k] = x[k] - x0[k];
            Quat q0(x0[6], x0[3], x0[4], x0[5]), q(x[6], x[3], x[4], x[5]);
            Quat dq = inverse(q0) * q;                 // positify() is the identity (utility.h:54-61)
            V3 v = 2.0 * dq.vec();
            if (!(dq.w >= 0)) v = -v;                  // :372-375
            for (int k = 0; k < 3; k++) dx[idx + 3 + k] = v[k];
        }
    }
    for (int i = 0; i < n; i++) {
        double s = info->linearized_residuals[i];
        for (int k = 0; k < n; k++) s += info->linearized_jacobians(i, k) * dx[k];
        residuals[i] = s;
    }
    if (jacobians) {
        for (size_t i = 0; i < info->keep_block_size.size(); i++) {
            if (!jacobians[i]) continue;
            int size = info->keep_block_size[i], local = (size == 7 ? 6 : size);
            int idx = info->keep_block_idx[i] - m;
            double *J = jacobians[i];
            for (int r = 0; r < n; r++) {
                for (int c = 0; c < size; c++) J[r * size + c] = 0.0;
                for (int c = 0; c < local; c++) J[r * size + c] = info->linearized_jacobians(r, idx + c);
            }
        }
    }
    return true;
}

}  // namespace oracle
