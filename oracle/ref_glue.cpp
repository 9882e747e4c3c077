// oracle/ref_glue.cpp -- TEST INFRASTRUCTURE.
//
// Thin extern "C" surface over the reference's OWN factor classes, compiled from the sources where they lie under
// /root/reference/src (see oracle/Makefile target `ref`) against the header shims of oracle/shim (Eigen / Ceres
// interfaces / ROS macros are not in this image).  It defines the configuration globals that parameters.cpp would
// load from the yaml (parameters.cpp needs OpenCV's FileStorage and is not compiled) and forwards to
//   Projection{TwoFrameOneCam,TwoFrameTwoCam,OneFrameTwoCam}Factor::Evaluate, IMULegFactor::Evaluate,
//   IMULegIntegrationBase::{push_back,...}, A1Kinematics::*, PoseLocalParameterization::Plus,
//   MarginalizationFactor::Evaluate.
// No reference source is copied into this repository; the output (oracle/_ref/libcerberus_ref.so) is git-ignored.
#include "../include/cerberus_b200.h"
#include "factor/projectionTwoFrameOneCamFactor.h"
#include "factor/projectionTwoFrameTwoCamFactor.h"
#include "factor/projectionOneFrameTwoCamFactor.h"
#include "factor/imu_leg_factor.h"
#include "factor/imu_leg_integration_base.h"
#include "factor/integration_base.h"
#include "factor/imu_factor.h"
#include "factor/marginalization_factor.h"
#include "factor/pose_local_parameterization.h"
#include "legKinematics/A1Kinematics.h"
#include "featureTracker/feature_manager.h"
#include "utils/utility.h"
#include <cstring>

// ---- globals of src/utils/parameters.cpp that the compiled objects reference --------------------------------
double ACC_N, ACC_N_Z, ACC_W, GYR_N, GYR_W;
Eigen::Vector3d G{0.0, 0.0, 9.805};
int CONTACT_SENSOR_TYPE;
double PHI_N, DPHI_N, RHO_C_N, RHO_NC_N;
double INIT_DEPTH = 5.0, MIN_PARALLAX = 10.0 / 460.0;      // parameters.cpp:250 (INIT_DEPTH), yaml keyframe_parallax / FOCAL_LENGTH
int NUM_OF_CAM = 2, STEREO = 1;
double V_N_MIN_XY, V_N_MIN_Z, V_N_MIN, V_N_MAX, V_N_FORCE_THRES_RATIO, V_N_TERM1_STEEP, V_N_TERM2_VAR_RESCALE, V_N_TERM3_DISTANCE_RESCALE;

namespace {
Eigen::Vector3d v3(const double *p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
struct LegCfg { std::vector<Eigen::VectorXd> rho_fix_list; Eigen::Vector3d p_br; Eigen::Matrix3d R_br; };
LegCfg g_leg;

IMULegIntegrationBase *make_integrator(const double *acc0, const double *gyr0, const double *phi0, const double *dphi0, const double *c0,
                                       const double *ba, const double *bg, const double *rho) {
    Vector_dof phi, dphi; Vector_leg c; Vector_rho r;
    for (int k = 0; k < 12; k++) { phi(k) = phi0[k]; dphi(k) = dphi0[k]; }
    for (int k = 0; k < 4; k++) { c(k) = c0[k]; r(k) = rho[k]; }
    return new IMULegIntegrationBase(v3(acc0), v3(gyr0), phi, dphi, c, v3(ba), v3(bg), r, g_leg.rho_fix_list, g_leg.p_br, g_leg.R_br);
}
void load_preint(IMULegIntegrationBase &p, const CerbIMULegPreint &q) {
    p.sum_dt = q.sum_dt; p.delta_p = v3(q.delta_p); p.delta_v = v3(q.delta_v);
    p.delta_q = Eigen::Quaterniond(q.delta_q[3], q.delta_q[0], q.delta_q[1], q.delta_q[2]);
    for (int k = 0; k < 4; k++) { p.delta_epsilon[k] = v3(q.delta_epsilon + 3 * k); p.linearized_rho(k) = q.linearized_rho[k]; }
    p.linearized_ba = v3(q.linearized_ba); p.linearized_bg = v3(q.linearized_bg);
    std::memcpy(p.jacobian.data(), q.jacobian, sizeof(q.jacobian));        // both column-major 31x31
    std::memcpy(p.covariance.data(), q.covariance, sizeof(q.covariance));
}
void store_preint(const IMULegIntegrationBase &p, CerbIMULegPreint &q) {
    q.sum_dt = p.sum_dt;
    for (int k = 0; k < 3; k++) { q.delta_p[k] = p.delta_p(k); q.delta_v[k] = p.delta_v(k); q.linearized_ba[k] = p.linearized_ba(k); q.linearized_bg[k] = p.linearized_bg(k); }
    q.delta_q[0] = p.delta_q.x(); q.delta_q[1] = p.delta_q.y(); q.delta_q[2] = p.delta_q.z(); q.delta_q[3] = p.delta_q.w();
    for (int j = 0; j < 4; j++) { for (int k = 0; k < 3; k++) q.delta_epsilon[3 * j + k] = p.delta_epsilon[j](k); q.linearized_rho[j] = p.linearized_rho(j); }
    std::memcpy(q.jacobian, p.jacobian.data(), sizeof(q.jacobian));
    std::memcpy(q.covariance, p.covariance.data(), sizeof(q.covariance));
}
const double kZero12[12] = {0}, kZero4[4] = {0}, kZero3[3] = {0};
}  // namespace

extern "C" {

void ref_set_globals(const CerbPreintConfig *c, const double *g, double visual_sqrt_info) {
    ACC_N = c->acc_n; ACC_N_Z = c->acc_n_z; GYR_N = c->gyr_n; ACC_W = c->acc_w; GYR_W = c->gyr_w; PHI_N = c->phi_n; DPHI_N = c->dphi_n;
    RHO_C_N = c->rho_c_n; RHO_NC_N = c->rho_nc_n; V_N_MIN_XY = c->v_n_min_xy; V_N_MIN_Z = c->v_n_min_z; V_N_MIN = c->v_n_min; V_N_MAX = c->v_n_max;
    V_N_FORCE_THRES_RATIO = c->v_n_force_thres_ratio; V_N_TERM1_STEEP = c->v_n_term1_steep; V_N_TERM2_VAR_RESCALE = c->v_n_term2_var_rescale;
    V_N_TERM3_DISTANCE_RESCALE = c->v_n_term3_distance_rescale; CONTACT_SENSOR_TYPE = c->contact_sensor_type;
    G = v3(g);
    // estimator.cpp:124-126
    ProjectionTwoFrameOneCamFactor::sqrt_info = visual_sqrt_info * Eigen::Matrix2d::Identity();
    ProjectionTwoFrameTwoCamFactor::sqrt_info = visual_sqrt_info * Eigen::Matrix2d::Identity();
    ProjectionOneFrameTwoCamFactor::sqrt_info = visual_sqrt_info * Eigen::Matrix2d::Identity();
    g_leg.rho_fix_list.clear();
    for (int l = 0; l < 4; l++) { Eigen::VectorXd f(RHO_FIX_SIZE); f << c->rho_fix[l][0], c->rho_fix[l][1], c->rho_fix[l][2], c->rho_fix[l][3]; g_leg.rho_fix_list.push_back(f); }
    g_leg.p_br = v3(c->p_br);
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) g_leg.R_br(r, k) = c->R_br[3 * r + k];
}

int ref_eval_projection(int kind, int n, const double *pose_i, const double *pose_j, const double *ex0, const double *ex1, const double *inv_dep,
                        const double *td, const double *pts_i, const double *pts_j, const double *vel_i, const double *vel_j, const double *td_i,
                        const double *td_j, double *residuals, double *jacobians) {
    for (int k = 0; k < n; k++) {
        Eigen::Vector3d pi = v3(pts_i + 3 * k), pj = v3(pts_j + 3 * k);
        Eigen::Vector2d vi(vel_i[2 * k], vel_i[2 * k + 1]), vj(vel_j[2 * k], vel_j[2 * k + 1]);
        double r[2]; double *J[6]; const double *p[6];
        if (kind == CERB_PROJ_TWO_FRAME_ONE_CAM) {
            ProjectionTwoFrameOneCamFactor f(pi, pj, vi, vj, td_i[k], td_j[k]);
            double *b = jacobians ? jacobians + (size_t)k * 46 : nullptr;
            p[0] = pose_i + 7 * k; p[1] = pose_j + 7 * k; p[2] = ex0 + 7 * k; p[3] = inv_dep + k; p[4] = td + k;
            if (b) { J[0] = b; J[1] = b + 14; J[2] = b + 28; J[3] = b + 42; J[4] = b + 44; }
            f.Evaluate(p, r, b ? J : nullptr);
        } else if (kind == CERB_PROJ_TWO_FRAME_TWO_CAM) {
            ProjectionTwoFrameTwoCamFactor f(pi, pj, vi, vj, td_i[k], td_j[k]);
            double *b = jacobians ? jacobians + (size_t)k * 60 : nullptr;
            p[0] = pose_i + 7 * k; p[1] = pose_j + 7 * k; p[2] = ex0 + 7 * k; p[3] = ex1 + 7 * k; p[4] = inv_dep + k; p[5] = td + k;
            if (b) { J[0] = b; J[1] = b + 14; J[2] = b + 28; J[3] = b + 42; J[4] = b + 56; J[5] = b + 58; }
            f.Evaluate(p, r, b ? J : nullptr);
        } else {
            ProjectionOneFrameTwoCamFactor f(pi, pj, vi, vj, td_i[k], td_j[k]);
            double *b = jacobians ? jacobians + (size_t)k * 32 : nullptr;
            p[0] = ex0 + 7 * k; p[1] = ex1 + 7 * k; p[2] = inv_dep + k; p[3] = td + k;
            if (b) { J[0] = b; J[1] = b + 14; J[2] = b + 28; J[3] = b + 30; }
            f.Evaluate(p, r, b ? J : nullptr);
        }
        if (residuals) { residuals[2 * k] = r[0]; residuals[2 * k + 1] = r[1]; }
    }
    return 0;
}

int ref_eval_imu_leg(int n, const CerbIMULegPreint *preint, const double *params, double *residuals, double *jacobians, double *sqrt_info) {
    for (int k = 0; k < n; k++) {
        IMULegIntegrationBase *pre = make_integrator(kZero3, kZero3, kZero12, kZero12, kZero4, preint[k].linearized_ba, preint[k].linearized_bg, preint[k].linearized_rho);
        load_preint(*pre, preint[k]);
        IMULegFactor f(pre);
        const double *q = params + (size_t)k * 40;
        const double *p[6] = {q, q + 7, q + 16, q + 20, q + 27, q + 36};
        double r[31]; double *J[6];
        double *b = jacobians ? jacobians + (size_t)k * 31 * 40 : nullptr;
        if (b) { J[0] = b; J[1] = b + 31 * 7; J[2] = b + 31 * 16; J[3] = b + 31 * 20; J[4] = b + 31 * 27; J[5] = b + 31 * 36; }
        f.Evaluate(p, r, b ? J : nullptr);
        if (residuals) for (int i = 0; i < 31; i++) residuals[(size_t)k * 31 + i] = r[i];
        if (sqrt_info) {   // the statement of imu_leg_factor.cpp:197-198 on the reference's own covariance member
            Eigen::Matrix<double, 31, 31> si = Eigen::LLT<Eigen::Matrix<double, 31, 31>>(pre->covariance.inverse()).matrixL().transpose();
            for (int i = 0; i < 31; i++) for (int j = 0; j < 31; j++) sqrt_info[(size_t)k * 961 + i * 31 + j] = si(i, j);
        }
        delete pre;
    }
    return 0;
}

int ref_preintegrate(int n, const CerbPreintJob *jobs, CerbIMULegPreint *out) {
    for (int k = 0; k < n; k++) {
        const CerbPreintJob &j = jobs[k];
        IMULegIntegrationBase *pre = make_integrator(j.acc_0, j.gyr_0, j.phi_0, j.dphi_0, j.c_0, j.linearized_ba, j.linearized_bg, j.linearized_rho);
        for (int s = 0; s < j.n_samples; s++) {
            const CerbIMULegSample &m = j.samples[s];
            Vector_dof phi, dphi; Vector_leg c;
            for (int t = 0; t < 12; t++) { phi(t) = m.phi[t]; dphi(t) = m.dphi[t]; }
            for (int t = 0; t < 4; t++) c(t) = m.c[t];
            pre->push_back(m.dt, v3(m.acc), v3(m.gyr), phi, dphi, c);
        }
        store_preint(*pre, out[k]);
        delete pre;
    }
    return 0;
}

int ref_a1_kinematics(int n, const double *q, const double *rho_opt, const double *rho_fix, double *fk, double *jac, double *dfk_drho, double *dJ_dq, double *dJ_drho) {
    A1Kinematics kin;
    for (int k = 0; k < n; k++) {
        Eigen::Vector3d qq = v3(q + 3 * k);
        Eigen::VectorXd ro(1), rf(4); ro << rho_opt[k]; rf << rho_fix[4 * k], rho_fix[4 * k + 1], rho_fix[4 * k + 2], rho_fix[4 * k + 3];
        if (fk) { Eigen::Vector3d o = kin.fk(qq, ro, rf); for (int t = 0; t < 3; t++) fk[3 * k + t] = o(t); }
        if (jac) { Eigen::Matrix3d o = kin.jac(qq, ro, rf); std::memcpy(jac + 9 * k, o.data(), 72); }
        if (dfk_drho) { Eigen::Matrix<double, 3, RHO_OPT_SIZE> o = kin.dfk_drho(qq, ro, rf); std::memcpy(dfk_drho + 3 * k, o.data(), 24); }
        if (dJ_dq) { Eigen::Matrix<double, 9, 3> o = kin.dJ_dq(qq, ro, rf); std::memcpy(dJ_dq + 27 * k, o.data(), 216); }
        if (dJ_drho) { Eigen::Matrix<double, 9, RHO_OPT_SIZE> o = kin.dJ_drho(qq, ro, rf); std::memcpy(dJ_drho + 9 * k, o.data(), 72); }
    }
    return 0;
}

int ref_pose_plus(const double *x, const double *delta, double *out) {
    PoseLocalParameterization p;
    static_cast<const ceres::LocalParameterization &>(p).Plus(x, delta, out);   // Plus is private in the subclass, public in the interface
    return 0;
}

// MarginalizationFactor::Evaluate of the reference on a prior given in the ABI shape
int ref_eval_prior(const CerbPrior *prior, const CerbWindowState *state, double *residuals, double *jacobians) {
    MarginalizationInfo *mi = new MarginalizationInfo();
    mi->n = prior->n; mi->m = 0;
    mi->linearized_jacobians = Eigen::MatrixXd(prior->n, prior->n);
    std::memcpy(mi->linearized_jacobians.data(), prior->linearized_jacobians, sizeof(double) * prior->n * prior->n);
    mi->linearized_residuals = Eigen::VectorXd(prior->n);
    for (int i = 0; i < prior->n; i++) mi->linearized_residuals(i) = prior->linearized_residuals[i];
    std::vector<const double *> params; std::vector<double *> J; size_t off = 0;
    CerbWindowState st = *state;
    for (int b = 0; b < prior->num_blocks; b++) {
        const int kind = prior->block_kind[b], idx = prior->block_index[b];
        const int size = (kind == CERB_BLOCK_POSE || kind == CERB_BLOCK_EX_POSE) ? 7 : (kind == CERB_BLOCK_SPEEDBIAS ? 9 : (kind == CERB_BLOCK_LEGBIAS ? 4 : 1));
        mi->keep_block_size.push_back(size); mi->keep_block_idx.push_back(prior->block_col[b]);
        mi->keep_block_data.push_back(const_cast<double *>(prior->block_x0[b]));
        const double *p = kind == CERB_BLOCK_POSE ? st.para_Pose[idx] : kind == CERB_BLOCK_SPEEDBIAS ? st.para_SpeedBias[idx] : kind == CERB_BLOCK_LEGBIAS ? st.para_LegBias[idx]
                        : kind == CERB_BLOCK_EX_POSE ? st.para_Ex_Pose[idx] : st.para_Td;
        params.push_back(p);
        J.push_back(jacobians ? jacobians + off : nullptr); off += (size_t)prior->n * size;
    }
    {
        MarginalizationFactor f(mi);
        f.Evaluate(params.data(), residuals, jacobians ? J.data() : nullptr);
    }
    mi->keep_block_data.clear();     // borrowed pointers
    delete mi;
    return 0;
}

// Marginalization half of Estimator::optimization() with the reference's OWN classes (ResidualBlockInfo,
// MarginalizationInfo::{addResidualBlockInfo, preMarginalize, marginalize, getParameterBlocks}, the factor classes,
// ceres::HuberLoss of the shim): the statements of estimator.cpp:1248-1376 (MARGIN_OLD) / :1377-1455 (SECOND_NEW)
// re-issued on the ABI structs.  Same output convention as oracle_marginalize (block order = the reference's
// unordered_map order, i.e. arbitrary).
int ref_marginalize(const CerbWindowDesc *desc, const CerbWindowState *state_in, int margin_old, CerbPrior *out, double *J_out, double *r_out) {
    static double para_Pose[CERB_NUM_FRAMES][7], para_SpeedBias[CERB_NUM_FRAMES][9], para_LegBias[CERB_NUM_FRAMES][4], para_Ex_Pose[2][7], para_Td[1][1];
    static double para_Feature[CERB_NUM_OF_F][1];
    std::memcpy(para_Pose, state_in->para_Pose, sizeof(para_Pose)); std::memcpy(para_SpeedBias, state_in->para_SpeedBias, sizeof(para_SpeedBias));
    std::memcpy(para_LegBias, state_in->para_LegBias, sizeof(para_LegBias)); std::memcpy(para_Ex_Pose, state_in->para_Ex_Pose, sizeof(para_Ex_Pose));
    para_Td[0][0] = state_in->para_Td[0];
    for (int f = 0; f < desc->n_features; f++) para_Feature[f][0] = state_in->para_Feature[f];
    auto block_ptr = [&](int kind, int idx) -> double * {
        return kind == CERB_BLOCK_POSE ? para_Pose[idx] : kind == CERB_BLOCK_SPEEDBIAS ? para_SpeedBias[idx] : kind == CERB_BLOCK_LEGBIAS ? para_LegBias[idx]
             : kind == CERB_BLOCK_EX_POSE ? para_Ex_Pose[idx] : para_Td[0]; };
    std::memset(out, 0, sizeof(*out));
    ceres::LossFunction *loss_function = new ceres::HuberLoss(1.0);
    MarginalizationInfo *last = nullptr; std::vector<double *> last_blocks;
    if (desc->prior.valid) {
        const CerbPrior &pr = desc->prior;
        last = new MarginalizationInfo(); last->n = pr.n; last->m = 0;
        last->linearized_jacobians = Eigen::MatrixXd(pr.n, pr.n);
        std::memcpy(last->linearized_jacobians.data(), pr.linearized_jacobians, sizeof(double) * pr.n * pr.n);
        last->linearized_residuals = Eigen::VectorXd(pr.n);
        for (int i = 0; i < pr.n; i++) last->linearized_residuals(i) = pr.linearized_residuals[i];
        for (int b = 0; b < pr.num_blocks; b++) {
            const int kind = pr.block_kind[b];
            const int size = (kind == CERB_BLOCK_POSE || kind == CERB_BLOCK_EX_POSE) ? 7 : (kind == CERB_BLOCK_SPEEDBIAS ? 9 : (kind == CERB_BLOCK_LEGBIAS ? 4 : 1));
            last->keep_block_size.push_back(size); last->keep_block_idx.push_back(pr.block_col[b]);
            double *copy = new double[size]; std::memcpy(copy, pr.block_x0[b], sizeof(double) * size);
            last->keep_block_data.push_back(copy);
            last_blocks.push_back(block_ptr(kind, pr.block_index[b]));
        }
    }
    MarginalizationInfo *mi = new MarginalizationInfo();
    std::vector<IMULegIntegrationBase *> keep_pre;
    if (margin_old) {
        if (last) {
            std::vector<int> drop_set;
            for (int i = 0; i < (int)last_blocks.size(); i++)
                if (last_blocks[i] == para_Pose[0] || last_blocks[i] == para_SpeedBias[0] || last_blocks[i] == para_LegBias[0]) drop_set.push_back(i);
            mi->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(last), NULL, last_blocks, drop_set));
        }
        if (desc->preint[0].sum_dt < 10.0) {
            IMULegIntegrationBase *pre = make_integrator(kZero3, kZero3, kZero12, kZero12, kZero4, desc->preint[0].linearized_ba, desc->preint[0].linearized_bg, desc->preint[0].linearized_rho);
            load_preint(*pre, desc->preint[0]); keep_pre.push_back(pre);
            mi->addResidualBlockInfo(new ResidualBlockInfo(new IMULegFactor(pre), NULL,
                std::vector<double *>{para_Pose[0], para_SpeedBias[0], para_LegBias[0], para_Pose[1], para_SpeedBias[1], para_LegBias[1]}, std::vector<int>{0, 1, 2}));
        }
        for (int fi = 0; fi < desc->n_features; fi++) {
            const CerbFeature &ft = desc->features[fi];
            if (ft.start_frame != 0) continue;
            const CerbObservation &o0 = desc->obs[ft.obs_offset];
            Eigen::Vector3d pts_i(o0.point[0], o0.point[1], 1.0); Eigen::Vector2d vel_i(o0.velocity[0], o0.velocity[1]);
            for (int k = 0; k < ft.n_obs; k++) {
                const CerbObservation &o = desc->obs[ft.obs_offset + k];
                const int imu_i = 0, imu_j = k;
                if (imu_i != imu_j) {
                    Eigen::Vector3d pts_j(o.point[0], o.point[1], 1.0);
                    auto *f_td = new ProjectionTwoFrameOneCamFactor(pts_i, pts_j, vel_i, Eigen::Vector2d(o.velocity[0], o.velocity[1]), o0.cur_td, o.cur_td);
                    mi->addResidualBlockInfo(new ResidualBlockInfo(f_td, loss_function,
                        std::vector<double *>{para_Pose[imu_i], para_Pose[imu_j], para_Ex_Pose[0], para_Feature[fi], para_Td[0]}, std::vector<int>{0, 3}));
                }
                if (o.is_stereo) {
                    Eigen::Vector3d pts_j_right(o.pointRight[0], o.pointRight[1], 1.0); Eigen::Vector2d vr(o.velocityRight[0], o.velocityRight[1]);
                    if (imu_i != imu_j) {
                        auto *f = new ProjectionTwoFrameTwoCamFactor(pts_i, pts_j_right, vel_i, vr, o0.cur_td, o.cur_td);
                        mi->addResidualBlockInfo(new ResidualBlockInfo(f, loss_function,
                            std::vector<double *>{para_Pose[imu_i], para_Pose[imu_j], para_Ex_Pose[0], para_Ex_Pose[1], para_Feature[fi], para_Td[0]}, std::vector<int>{0, 4}));
                    } else {
                        auto *f = new ProjectionOneFrameTwoCamFactor(pts_i, pts_j_right, vel_i, vr, o0.cur_td, o.cur_td);
                        mi->addResidualBlockInfo(new ResidualBlockInfo(f, loss_function,
                            std::vector<double *>{para_Ex_Pose[0], para_Ex_Pose[1], para_Feature[fi], para_Td[0]}, std::vector<int>{2}));
                    }
                }
            }
        }
    } else {
        bool has = false;
        for (double *p : last_blocks) if (p == para_Pose[CERB_WINDOW_SIZE - 1]) has = true;
        if (!has) { *out = desc->prior; return 0; }
        std::vector<int> drop_set;
        for (int i = 0; i < (int)last_blocks.size(); i++) if (last_blocks[i] == para_Pose[CERB_WINDOW_SIZE - 1]) drop_set.push_back(i);
        mi->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(last), NULL, last_blocks, drop_set));
    }
    mi->preMarginalize();
    mi->marginalize();
    if (!mi->valid) { out->valid = 0; return 0; }
    std::unordered_map<long, double *> addr_shift;
    if (margin_old) {
        for (int i = 1; i <= CERB_WINDOW_SIZE; i++) {
            addr_shift[reinterpret_cast<long>(para_Pose[i])] = para_Pose[i - 1];
            addr_shift[reinterpret_cast<long>(para_SpeedBias[i])] = para_SpeedBias[i - 1];
            addr_shift[reinterpret_cast<long>(para_LegBias[i])] = para_LegBias[i - 1];
        }
    } else {
        for (int i = 0; i <= CERB_WINDOW_SIZE; i++) {
            if (i == CERB_WINDOW_SIZE - 1) continue;
            const int t = (i == CERB_WINDOW_SIZE) ? i - 1 : i;
            addr_shift[reinterpret_cast<long>(para_Pose[i])] = para_Pose[t];
            addr_shift[reinterpret_cast<long>(para_SpeedBias[i])] = para_SpeedBias[t];
            addr_shift[reinterpret_cast<long>(para_LegBias[i])] = para_LegBias[t];
        }
    }
    for (int i = 0; i < 2; i++) addr_shift[reinterpret_cast<long>(para_Ex_Pose[i])] = para_Ex_Pose[i];
    addr_shift[reinterpret_cast<long>(para_Td[0])] = para_Td[0];
    std::vector<double *> parameter_blocks = mi->getParameterBlocks(addr_shift);
    out->valid = 1; out->n = mi->n; out->num_blocks = (int)parameter_blocks.size();
    for (int b = 0; b < out->num_blocks; b++) {
        double *p = parameter_blocks[b]; int kind = -1, index = 0;
        for (int i = 0; i < CERB_NUM_FRAMES; i++) {
            if (p == para_Pose[i]) { kind = CERB_BLOCK_POSE; index = i; }
            if (p == para_SpeedBias[i]) { kind = CERB_BLOCK_SPEEDBIAS; index = i; }
            if (p == para_LegBias[i]) { kind = CERB_BLOCK_LEGBIAS; index = i; }
        }
        for (int i = 0; i < 2; i++) if (p == para_Ex_Pose[i]) { kind = CERB_BLOCK_EX_POSE; index = i; }
        if (p == para_Td[0]) { kind = CERB_BLOCK_TD; index = 0; }
        out->block_kind[b] = kind; out->block_index[b] = index; out->block_col[b] = mi->keep_block_idx[b] - mi->m;
        for (int k = 0; k < mi->keep_block_size[b]; k++) out->block_x0[b][k] = mi->keep_block_data[b][k];
    }
    std::memcpy(J_out, mi->linearized_jacobians.data(), sizeof(double) * mi->n * mi->n);
    for (int i = 0; i < mi->n; i++) r_out[i] = mi->linearized_residuals(i);
    out->linearized_jacobians = J_out; out->linearized_residuals = r_out;
    // (objects are leaked on purpose: MarginalizationInfo's destructor owns the factors and `last`'s data; test helper only)
    return 0;
}

// IMUFactor::Evaluate / IntegrationBase (imu_factor.h, integration_base.h) -- the USE_LEG == 0 path
int ref_eval_imu(int n, const CerbIMUPreint *preint, const double *params, double *residuals, double *jacobians, double *sqrt_info) {
    for (int k = 0; k < n; k++) {
        const CerbIMUPreint &q = preint[k];
        IntegrationBase pre(v3(kZero3), v3(kZero3), v3(q.linearized_ba), v3(q.linearized_bg));
        pre.sum_dt = q.sum_dt; pre.delta_p = v3(q.delta_p); pre.delta_v = v3(q.delta_v);
        pre.delta_q = Eigen::Quaterniond(q.delta_q[3], q.delta_q[0], q.delta_q[1], q.delta_q[2]);
        std::memcpy(pre.jacobian.data(), q.jacobian, sizeof(q.jacobian)); std::memcpy(pre.covariance.data(), q.covariance, sizeof(q.covariance));
        IMUFactor f(&pre);
        const double *x = params + (size_t)k * 32;
        const double *p[4] = {x, x + 7, x + 16, x + 23};
        double r[15]; double *J[4];
        double *b = jacobians ? jacobians + (size_t)k * 15 * 32 : nullptr;
        if (b) { J[0] = b; J[1] = b + 15 * 7; J[2] = b + 15 * 16; J[3] = b + 15 * 23; }
        f.Evaluate(p, r, b ? J : nullptr);
        if (residuals) for (int i = 0; i < 15; i++) residuals[(size_t)k * 15 + i] = r[i];
        if (sqrt_info) {
            Eigen::Matrix<double, 15, 15> si = Eigen::LLT<Eigen::Matrix<double, 15, 15>>(pre.covariance.inverse()).matrixL().transpose();
            for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) sqrt_info[(size_t)k * 225 + i * 15 + j] = si(i, j);
        }
    }
    return 0;
}
int ref_preintegrate_imu(int n, const CerbPreintJob *jobs, CerbIMUPreint *out) {
    for (int k = 0; k < n; k++) {
        const CerbPreintJob &j = jobs[k];
        IntegrationBase pre(v3(j.acc_0), v3(j.gyr_0), v3(j.linearized_ba), v3(j.linearized_bg));
        for (int s = 0; s < j.n_samples; s++) pre.push_back(j.samples[s].dt, v3(j.samples[s].acc), v3(j.samples[s].gyr));
        CerbIMUPreint &q = out[k];
        q.sum_dt = pre.sum_dt;
        for (int t = 0; t < 3; t++) { q.delta_p[t] = pre.delta_p(t); q.delta_v[t] = pre.delta_v(t); q.linearized_ba[t] = pre.linearized_ba(t); q.linearized_bg[t] = pre.linearized_bg(t); }
        q.delta_q[0] = pre.delta_q.x(); q.delta_q[1] = pre.delta_q.y(); q.delta_q[2] = pre.delta_q.z(); q.delta_q[3] = pre.delta_q.w();
        std::memcpy(q.jacobian, pre.jacobian.data(), sizeof(q.jacobian)); std::memcpy(q.covariance, pre.covariance.data(), sizeof(q.covariance));
    }
    return 0;
}

// ---- rows a2 / n3: the host-side steps either side of the solve, on the reference's own code --------------------------------------
void ref_set_eigen_mode(int mode) { Eigen::shim_eig_mode() = mode; }    // 0: tridiagonal QR (Eigen's algorithm), 1: cyclic Jacobi

// Estimator::double2vector, estimator.cpp:903-957 (USE_IMU branch), re-issued statement by statement on Utility::R2ypr / ypr2R
// (utils/utility.h:85-125, compiled from the reference) and the quaternion -> rotation conversions of the Eigen shim.
// `before` holds the para_* arrays written by vector2double() (Rs[0] = its quaternion as a matrix, Ps[0] = its position).
void ref_double2vector(const CerbWindowState *before, const CerbWindowState *after, double *Ps_out, double *Rs_out, double *Vs_out) {
    using namespace Eigen;
    const double (*para_Pose)[7] = after->para_Pose; const double (*para_SpeedBias)[9] = after->para_SpeedBias;
    Matrix3d Rs0 = Quaterniond(before->para_Pose[0][6], before->para_Pose[0][3], before->para_Pose[0][4], before->para_Pose[0][5]).toRotationMatrix();
    Vector3d origin_R0 = Utility::R2ypr(Rs0);
    Vector3d origin_P0(before->para_Pose[0][0], before->para_Pose[0][1], before->para_Pose[0][2]);
    Vector3d origin_R00 = Utility::R2ypr(Quaterniond(para_Pose[0][6], para_Pose[0][3], para_Pose[0][4], para_Pose[0][5]).toRotationMatrix());
    double y_diff = origin_R0.x() - origin_R00.x();
    Matrix3d rot_diff = Utility::ypr2R(Vector3d(y_diff, 0, 0));
    if (abs(abs(origin_R0.y()) - 90) < 1.0 || abs(abs(origin_R00.y()) - 90) < 1.0)
        rot_diff = Rs0 * Quaterniond(para_Pose[0][6], para_Pose[0][3], para_Pose[0][4], para_Pose[0][5]).toRotationMatrix().transpose();
    for (int i = 0; i <= WINDOW_SIZE; i++) {
        Matrix3d R = rot_diff * Quaterniond(para_Pose[i][6], para_Pose[i][3], para_Pose[i][4], para_Pose[i][5]).normalized().toRotationMatrix();
        Vector3d P = rot_diff * Vector3d(para_Pose[i][0] - para_Pose[0][0], para_Pose[i][1] - para_Pose[0][1], para_Pose[i][2] - para_Pose[0][2]) + origin_P0;
        Vector3d V = rot_diff * Vector3d(para_SpeedBias[i][0], para_SpeedBias[i][1], para_SpeedBias[i][2]);
        for (int k = 0; k < 3; k++) { Ps_out[3 * i + k] = P(k); Vs_out[3 * i + k] = V(k); for (int c = 0; c < 3; c++) Rs_out[9 * i + 3 * k + c] = R(k, c); }
    }
}

namespace {
// FeatureManager of the reference filled from the ABI descriptors (feature_id = index in the caller's array) + Rs / Ps / ric / tic
struct RefWindow {
    Eigen::Matrix3d Rs[WINDOW_SIZE + 1], ric[2]; Eigen::Vector3d Ps[WINDOW_SIZE + 1], tic[2];
    FeatureManager fm;
    RefWindow(const CerbWindowDesc *d, const CerbWindowState *st, bool depth_from_state) : fm(Rs) {
        for (int i = 0; i <= WINDOW_SIZE; i++) {
            Rs[i] = Eigen::Quaterniond(st->para_Pose[i][6], st->para_Pose[i][3], st->para_Pose[i][4], st->para_Pose[i][5]).normalized().toRotationMatrix();
            Ps[i] = Eigen::Vector3d(st->para_Pose[i][0], st->para_Pose[i][1], st->para_Pose[i][2]);
        }
        for (int c = 0; c < 2; c++) {
            ric[c] = Eigen::Quaterniond(st->para_Ex_Pose[c][6], st->para_Ex_Pose[c][3], st->para_Ex_Pose[c][4], st->para_Ex_Pose[c][5]).normalized().toRotationMatrix();
            tic[c] = Eigen::Vector3d(st->para_Ex_Pose[c][0], st->para_Ex_Pose[c][1], st->para_Ex_Pose[c][2]);
        }
        fm.setRic(ric);
        for (int f = 0; f < d->n_features; f++) {
            const CerbFeature &ft = d->features[f];
            fm.feature.push_back(FeaturePerId(f, ft.start_frame));
            for (int k = 0; k < ft.n_obs; k++) {
                const CerbObservation &o = d->obs[ft.obs_offset + k];
                Eigen::Matrix<double, 7, 1> p; p << o.point[0], o.point[1], 1.0, 0.0, 0.0, o.velocity[0], o.velocity[1];
                FeaturePerFrame fpf(p, o.cur_td);
                if (o.is_stereo) { Eigen::Matrix<double, 7, 1> q; q << o.pointRight[0], o.pointRight[1], 1.0, 0.0, 0.0, o.velocityRight[0], o.velocityRight[1]; fpf.rightObservation(q); }
                fm.feature.back().feature_per_frame.push_back(fpf);
            }
            if (depth_from_state) fm.feature.back().estimated_depth = 1.0 / st->para_Feature[f];        // FeatureManager::setDepth, feature_manager.cpp:142-160
        }
    }
};
// Estimator::reprojectionError, estimator.cpp:1729-1739 (a member of Estimator, which cannot be compiled here: statements re-issued)
double reprojectionError(Eigen::Matrix3d &Ri, Eigen::Vector3d &Pi, Eigen::Matrix3d &rici, Eigen::Vector3d &tici, Eigen::Matrix3d &Rj, Eigen::Vector3d &Pj,
                         Eigen::Matrix3d &ricj, Eigen::Vector3d &ticj, double depth, Eigen::Vector3d &uvi, Eigen::Vector3d &uvj) {
    Eigen::Vector3d pts_w = Ri * (rici * (depth * uvi) + tici) + Pi;
    Eigen::Vector3d pts_cj = ricj.transpose() * (Rj.transpose() * (pts_w - Pj) - ticj);
    Eigen::Vector2d residual = (pts_cj / pts_cj.z()).head<2>() - uvj.head<2>();
    double rx = residual.x(), ry = residual.y();
    return sqrt(rx * rx + ry * ry);
}
}  // namespace

// FeatureManager::triangulate (feature_manager.cpp:302-431) of the reference on the window: features with para_Feature <= 0 start
// untriangulated (estimated_depth = -1, FeaturePerId ctor), the others keep 1 / para_Feature.
int ref_triangulate(const CerbWindowDesc *d, const CerbWindowState *st, double init_depth, double *depth) {
    INIT_DEPTH = init_depth; STEREO = 1;
    RefWindow W(d, st, false);
    { int f = 0; for (auto &it : W.fm.feature) { if (st->para_Feature[f] > 0.0) it.estimated_depth = 1.0 / st->para_Feature[f]; f++; } }
    W.fm.triangulate(WINDOW_SIZE, W.Ps, W.Rs, W.tic, W.ric);
    for (auto &it : W.fm.feature) depth[it.feature_id] = it.estimated_depth;
    return 0;
}

// Estimator::slideWindowOld + FeatureManager::removeBackShiftDepth (estimator.cpp:1660-1677, feature_manager.cpp:450-488): back_R0 / back_P0
// are the marginalized frame 0, Rs[0] / Ps[0] after the slide are the old frame 1.
int ref_shift_depth(const CerbWindowDesc *d, const CerbWindowState *st, double init_depth, int *new_start, double *depth, int *keep) {
    INIT_DEPTH = init_depth;
    RefWindow W(d, st, true);
    Eigen::Matrix3d back_R0 = W.Rs[0]; Eigen::Vector3d back_P0 = W.Ps[0];
    Eigen::Matrix3d R0, R1; Eigen::Vector3d P0, P1;
    R0 = back_R0 * W.ric[0];
    R1 = W.Rs[1] * W.ric[0];
    P0 = back_P0 + back_R0 * W.tic[0];
    P1 = W.Ps[1] + W.Rs[1] * W.tic[0];
    for (int f = 0; f < d->n_features; f++) keep[f] = 0;
    W.fm.removeBackShiftDepth(R0, P0, R1, P1);
    for (auto &it : W.fm.feature) { keep[it.feature_id] = 1; new_start[it.feature_id] = it.start_frame; depth[it.feature_id] = it.estimated_depth; }
    return 0;
}

// Estimator::outliersRejection, estimator.cpp:1741-1798, statements re-issued on the reference's FeatureManager list; ave_err per feature
int ref_outlier_errors(const CerbWindowDesc *d, const CerbWindowState *st, double *ave_err_out) {
    STEREO = 1;
    RefWindow W(d, st, true);
    Eigen::Matrix3d *Rs = W.Rs, *ric = W.ric; Eigen::Vector3d *Ps = W.Ps, *tic = W.tic;
    for (auto &it_per_id : W.fm.feature) {
        double err = 0; int errCnt = 0;
        it_per_id.used_num = it_per_id.feature_per_frame.size();
        if (it_per_id.used_num < 4) continue;
        int imu_i = it_per_id.start_frame, imu_j = imu_i - 1;
        Eigen::Vector3d pts_i = it_per_id.feature_per_frame[0].point;
        double depth = it_per_id.estimated_depth;
        for (auto &it_per_frame : it_per_id.feature_per_frame) {
            imu_j++;
            if (imu_i != imu_j) {
                Eigen::Vector3d pts_j = it_per_frame.point;
                err += reprojectionError(Rs[imu_i], Ps[imu_i], ric[0], tic[0], Rs[imu_j], Ps[imu_j], ric[0], tic[0], depth, pts_i, pts_j); errCnt++;
            }
            if (STEREO && it_per_frame.is_stereo) {
                Eigen::Vector3d pts_j_right = it_per_frame.pointRight;
                err += reprojectionError(Rs[imu_i], Ps[imu_i], ric[0], tic[0], Rs[imu_j], Ps[imu_j], ric[1], tic[1], depth, pts_i, pts_j_right); errCnt++;
            }
        }
        ave_err_out[it_per_id.feature_id] = err / errCnt;
    }
    return 0;
}

}  // extern "C"
